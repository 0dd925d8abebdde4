// SURVEY.md 8(f) rows 3 and 4 on the device.
//
// f3 -- one SCF step without a host round-trip of matrices: getC / solveFockMatrix (src/HartreeFock.jl:39-56), getD
// (:296-302), getG / getF (:322-335), getE (:339-350), the residual F D S - S D F (:1298-1299) and the Gram matrices
// behind DIIS / EDIIS / ADIIS (:1273-1316) as kernels and FP64 GEMMs on the library's stream, `eigen` through
// cuSOLVER (linalg.cu).  The host keeps what is tiny: the m x m coefficient problem of the extrapolation and the
// stage / convergence logic.  The Fock build inside a step is the hot path itself (qbx_fock_device: digestion of the
// packed store + the all-reduce over the communicator's ranks).
//
// f4 -- changeOrbitalBasis (src/Integration/Interface.jl:376-407): (ij|kl) = sum (ab|cd) C_ai C_bj C_ck C_dl as four
// quarter transforms, each a (strided-batched) FP64 GEMM over the dense tensor on the device, and the alpha-beta Coulomb
// matrix of the two-coefficient method (getJab, :384-393).
#include <string.h>

#include "handle.h"
#include "linalg.h"

namespace {
__global__ void k_scale_cols(int n, const double *w, double *U)          // U[:, j] *= w[j]^(-1/4)  (X = V V^T, V = U w^-1/4)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < (int64_t)n * n) U[e] *= 1.0 / sqrt(sqrt(w[e / n]));
}
__global__ void k_sign_fix(int n, double *C)                             // column sign: first row non-negative (:53)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < (int64_t)n * n && C[(e / n) * n] < 0.0 && e % n != 0) C[e] = -C[e];
}
__global__ void k_sign_fix_row0(int n, double *C)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && C[(int64_t)j * n] < 0.0) C[(int64_t)j * n] = -C[(int64_t)j * n];
}
__global__ void k_axpby(int64_t n, double a, const double *x, double b, const double *y, double *out)   // out = a x + b y
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < n) out[e] = a * x[e] + (y ? b * y[e] : 0.0);
}
__global__ void k_accum(int64_t n, double a, const double *x, double *out)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < n) out[e] += a * x[e];
}
// out[0] += <x, y>; out[1] += <x - y, x - y>   (one launch, block reduction, fp64 atomics)
__global__ void k_dot2(int64_t n, const double *x, const double *y, double *out)
{
    __shared__ double r0[256], r1[256];
    double a = 0, b = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        a += x[e] * y[e];
        const double d = x[e] - y[e];
        b += d * d;
    }
    r0[threadIdx.x] = a; r1[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { r0[threadIdx.x] += r0[threadIdx.x + s]; r1[threadIdx.x] += r1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(out, r0[0]); atomicAdd(out + 1, r1[0]); }
}
inline unsigned grid1(int64_t n) { return (unsigned)((n + 255) / 256); }
}   // namespace

struct qbx_scf {
    qbx_basis *b = nullptr;
    int n = 0, cap = 0;
    double *S = nullptr, *H = nullptr, *X = nullptr;
    double *Fin = nullptr, *C = nullptr, *D = nullptr, *F = nullptr, *Dprev = nullptr;   // [2][n^2] each
    double *eps = nullptr;                                                                   // [2][n]
    double *R = nullptr;                                                                     // [2][n^2]: X^T (F D S - S D F) X
    double *DJ = nullptr, *G = nullptr, *t1 = nullptr, *t2 = nullptr;                        // n^2 (G: 2 n^2)
    double *hD = nullptr, *hF = nullptr, *hR = nullptr;                                      // history [cap][2][n^2]
    double *d_sc = nullptr;                                                                  // device scalars
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // step start, solved, Fock built, end; [4..5] bracket the first build of a :DD step
    double t_solve = 0, t_fock = 0, t_total = 0, nsteps = 0;                                 // device seconds since creation
};

static void scf_free(qbx_scf *s)
{
    for (double *p : {s->S, s->H, s->X, s->Fin, s->C, s->D, s->F, s->Dprev, s->eps, s->R, s->DJ, s->G, s->t1, s->t2, s->hD, s->hF, s->hR, s->d_sc})
        qbx_pool_free(p);
    for (cudaEvent_t e : s->ev) if (e) cudaEventDestroy(e);
    delete s;
}

extern "C" int qbx_scf_create(qbx_basis *b, const double *S, const double *Hcore, int history, qbx_scf **out)
{
    if (!b || !S || !Hcore || !out || history < 1 || history > 64) { qbx_set_error("qbx_scf_create: bad argument"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    cudaStream_t st = qbx_stream();
    qbx_scf *s = new qbx_scf;
    s->b = b; s->n = (int)b->nbf; s->cap = history;
    const size_t n2 = (size_t)s->n * s->n, B2 = n2 * sizeof(double);
    double **one[] = {&s->S, &s->H, &s->X, &s->DJ, &s->t1, &s->t2};
    double **two[] = {&s->Fin, &s->C, &s->D, &s->F, &s->Dprev, &s->R, &s->G};
    for (double **p : one) if (qbx_dmalloc(p, B2) != cudaSuccess) { scf_free(s); qbx_set_error("qbx_scf_create: out of device memory"); return QBX_ERR_NOMEM; }
    for (double **p : two) if (qbx_dmalloc(p, 2 * B2) != cudaSuccess) { scf_free(s); qbx_set_error("qbx_scf_create: out of device memory"); return QBX_ERR_NOMEM; }
    double **hist[] = {&s->hD, &s->hF, &s->hR};
    for (double **p : hist) if (qbx_dmalloc(p, 2 * B2 * history) != cudaSuccess) { scf_free(s); qbx_set_error("qbx_scf_create: out of device memory"); return QBX_ERR_NOMEM; }
    QBX_CUDA(qbx_dmalloc(&s->eps, 2 * s->n * sizeof(double)));
    QBX_CUDA(qbx_dmalloc(&s->d_sc, 64 * sizeof(double)));
    for (cudaEvent_t &e : s->ev) QBX_CUDA(cudaEventCreate(&e));
    QBX_CUDA(cudaMemcpyAsync(s->S, S, B2, cudaMemcpyHostToDevice, st));
    QBX_CUDA(cudaMemcpyAsync(s->H, Hcore, B2, cudaMemcpyHostToDevice, st));
    QBX_CUDA(cudaMemsetAsync(s->D, 0, 2 * B2, st));
    QBX_CUDA(cudaMemsetAsync(s->F, 0, 2 * B2, st));
    // X = S^(-1/2) (getOrthonormalization, HartreeFock.jl:39-42): S = U w U^T, V = U w^(-1/4), X = V V^T
    QBX_CUDA(cudaMemcpyAsync(s->t1, s->S, B2, cudaMemcpyDeviceToDevice, st));
    if ((rc = qbx_syevd(s->n, s->t1, s->eps, st))) { scf_free(s); return rc; }
    k_scale_cols<<<grid1(n2), 256, 0, st>>>(s->n, s->eps, s->t1);
    if ((rc = qbx_gemm(0, 1, s->n, s->n, s->n, 1.0, s->t1, s->n, s->t1, s->n, 0.0, s->X, s->n, st))) { scf_free(s); return rc; }
    QBX_CUDA(cudaStreamSynchronize(st));
    *out = s;
    return QBX_OK;
}

extern "C" int qbx_scf_destroy(qbx_scf *s)
{
    if (!s) return QBX_OK;
    if (qbx_ensure_init() == QBX_OK) cudaStreamSynchronize(qbx_stream());
    scf_free(s);
    return QBX_OK;
}

// which: 0 Fin (the matrix the next step diagonalises), 1 C
extern "C" int qbx_scf_set(qbx_scf *s, int which, int spin, const double *M)
{
    if (!s || !M || spin < 0 || spin > 1 || which < 0 || which > 1) { qbx_set_error("qbx_scf_set: bad argument"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    const size_t n2 = (size_t)s->n * s->n;
    QBX_CUDA(cudaMemcpyAsync((which == 0 ? s->Fin : s->C) + spin * n2, M, n2 * sizeof(double), cudaMemcpyHostToDevice, qbx_stream()));
    QBX_CUDA(cudaStreamSynchronize(qbx_stream()));
    return QBX_OK;
}

// which: 0 Fin, 1 C, 2 D, 3 F, 4 eps (n doubles), 5 X, 6 device seconds {eigen + transforms, Fock builds, whole steps, steps}
extern "C" int qbx_scf_get(qbx_scf *s, int which, int spin, double *M)
{
    if (!s || !M || spin < 0 || spin > 1 || which < 0 || which > 6) { qbx_set_error("qbx_scf_get: bad argument"); return QBX_ERR_ARG; }
    if (which == 6) { M[0] = s->t_solve; M[1] = s->t_fock; M[2] = s->t_total; M[3] = s->nsteps; return QBX_OK; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    const size_t n2 = (size_t)s->n * s->n;
    const double *src[] = {s->Fin + spin * n2, s->C + spin * n2, s->D + spin * n2, s->F + spin * n2, s->eps + spin * s->n, s->X};
    QBX_CUDA(cudaMemcpyAsync(M, src[which], (which == 4 ? (size_t)s->n : n2) * sizeof(double), cudaMemcpyDeviceToHost, qbx_stream()));
    QBX_CUDA(cudaStreamSynchronize(qbx_stream()));
    return QBX_OK;
}

// C, eps <- generalised eigenproblem of Fin through X (solveFockMatrix :46-56), sign-fixed columns
static int scf_solve(qbx_scf *s, int spin, cudaStream_t st)
{
    const int n = s->n;
    const size_t n2 = (size_t)n * n;
    int rc;
    if ((rc = qbx_gemm(1, 0, n, n, n, 1.0, s->X, n, s->Fin + spin * n2, n, 0.0, s->t1, n, st))) return rc;        // X^T F
    if ((rc = qbx_gemm(0, 0, n, n, n, 1.0, s->t1, n, s->X, n, 0.0, s->t2, n, st))) return rc;                      // (X^T F) X
    if ((rc = qbx_syevd(n, s->t2, s->eps + spin * n, st))) return rc;
    if ((rc = qbx_gemm(0, 0, n, n, n, 1.0, s->X, n, s->t2, n, 0.0, s->C + spin * n2, n, st))) return rc;           // C = X C'
    k_sign_fix<<<grid1(n2), 256, 0, st>>>(n, s->C + spin * n2);
    k_sign_fix_row0<<<grid1(n), 256, 0, st>>>(n, s->C + spin * n2);
    return QBX_OK;
}

// G(D) for the current densities in s->D -> s->G  (getG :322-327): RHF DJ = 2 D, DK = D; UHF DJ = Da + Db, DK = (Da, Db)
static int scf_fock(qbx_scf *s, int nspin, const double *D, cudaStream_t st)
{
    const size_t n2 = (size_t)s->n * s->n;
    if (nspin == 1) k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 2.0, D, 0.0, nullptr, s->DJ);
    else k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 1.0, D, 1.0, D + n2, s->DJ);
    std::lock_guard<std::mutex> lk(s->b->mu);
    return qbx_fock_device(s->b, nspin, s->DJ, D, s->G, st);
}

/* One getCDFE (HartreeFock.jl:392-403) for nspin sectors, entirely on the device.
 *   from_coeff = 0: diagonalise Fin; 1: keep the coefficients already in C (initial guesses given as C)
 *   damp > 0 (the :DD step, :1245-1270): the density fed to the Fock build is (1 - damp) D_new + damp D_previous, and
 *            the matrix diagonalised is the CURRENT F
 * out[0..1] E per spin (getE), out[2] mean over spins of RMS(F D S - S D F), out[3] RMS of the change of the total density */
extern "C" int qbx_scf_step(qbx_scf *s, int nspin, const int *nocc, int from_coeff, double damp, double *out)
{
    if (!s || !nocc || !out || nspin < 1 || nspin > 2) { qbx_set_error("qbx_scf_step: bad argument"); return QBX_ERR_ARG; }
    for (int k = 0; k < nspin; ++k)
        if (nocc[k] < 0 || nocc[k] > s->n) { qbx_set_error("qbx_scf_step: occupation out of range"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    cudaStream_t st = qbx_stream();
    const int n = s->n;
    const size_t n2 = (size_t)n * n;
    QBX_CUDA(cudaEventRecord(s->ev[0], st));
    QBX_CUDA(cudaMemcpyAsync(s->Dprev, s->D, 2 * n2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (damp > 0.0) {
        // :DD = two Fock builds: F_in = Hcore + G((1 - damp) D(C(F)) + damp D), then the ordinary step on F_in
        for (int k = 0; k < nspin; ++k) {
            QBX_CUDA(cudaMemcpyAsync(s->Fin + k * n2, s->F + k * n2, n2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
            if ((rc = scf_solve(s, k, st))) return rc;
            if (nocc[k] > 0) { if ((rc = qbx_gemm(0, 1, n, n, nocc[k], 1.0 - damp, s->C + k * n2, n, s->C + k * n2, n, damp, s->D + k * n2, n, st))) return rc; }
            else k_axpby<<<grid1(n2), 256, 0, st>>>(n2, damp, s->D + k * n2, 0.0, nullptr, s->D + k * n2);
        }
        QBX_CUDA(cudaEventRecord(s->ev[4], st));
        if ((rc = scf_fock(s, nspin, s->D, st))) return rc;
        QBX_CUDA(cudaEventRecord(s->ev[5], st));
        for (int k = 0; k < nspin; ++k) k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 1.0, s->H, 1.0, s->G + k * n2, s->Fin + k * n2);
    }
    for (int k = 0; k < nspin; ++k) {
        if (!from_coeff && (rc = scf_solve(s, k, st))) return rc;
        if (nocc[k] > 0) { if ((rc = qbx_gemm(0, 1, n, n, nocc[k], 1.0, s->C + k * n2, n, s->C + k * n2, n, 0.0, s->D + k * n2, n, st))) return rc; }
        else QBX_CUDA(cudaMemsetAsync(s->D + k * n2, 0, n2 * sizeof(double), st));
    }
    QBX_CUDA(cudaEventRecord(s->ev[1], st));
    if ((rc = scf_fock(s, nspin, s->D, st))) return rc;
    QBX_CUDA(cudaEventRecord(s->ev[2], st));
    QBX_CUDA(cudaMemsetAsync(s->d_sc, 0, 64 * sizeof(double), st));
    for (int k = 0; k < nspin; ++k) {
        double *F = s->F + k * n2, *D = s->D + k * n2;
        k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 1.0, s->H, 1.0, s->G + k * n2, F);                      // F = Hcore + G
        k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 1.0, s->H, 1.0, F, s->t1);                              // Hcore + F
        k_dot2<<<148, 256, 0, st>>>(n2, D, s->t1, s->d_sc + 4 * k);                                    // [4k] = <D, Hcore + F>
        // R = F D S - S D F; kept X-transformed for the DIIS Gram matrix
        if ((rc = qbx_gemm(0, 0, n, n, n, 1.0, D, n, s->S, n, 0.0, s->t1, n, st))) return rc;          // D S
        if ((rc = qbx_gemm(0, 0, n, n, n, 1.0, F, n, s->t1, n, 0.0, s->t2, n, st))) return rc;         // F D S
        if ((rc = qbx_gemm(1, 1, n, n, n, -1.0, s->t1, n, F, n, 1.0, s->t2, n, st))) return rc;        // - (D S)^T F^T = - S D F
        QBX_CUDA(cudaMemsetAsync(s->t1, 0, n2 * sizeof(double), st));
        k_dot2<<<148, 256, 0, st>>>(n2, s->t2, s->t1, s->d_sc + 8 + 4 * k);                            // [8 + 4k + 1] = |R|^2
        if ((rc = qbx_gemm(1, 0, n, n, n, 1.0, s->X, n, s->t2, n, 0.0, s->t1, n, st))) return rc;
        if ((rc = qbx_gemm(0, 0, n, n, n, 1.0, s->t1, n, s->X, n, 0.0, s->R + k * n2, n, st))) return rc;
    }
    // change of the total density (2 D for RHF, Da + Db for UHF)
    if (nspin == 1) { k_dot2<<<148, 256, 0, st>>>(n2, s->D, s->Dprev, s->d_sc + 16); }
    else {
        k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 1.0, s->D, 1.0, s->D + n2, s->t1);
        k_axpby<<<grid1(n2), 256, 0, st>>>(n2, 1.0, s->Dprev, 1.0, s->Dprev + n2, s->t2);
        k_dot2<<<148, 256, 0, st>>>(n2, s->t1, s->t2, s->d_sc + 16);
    }
    QBX_CUDA(cudaGetLastError());
    double *h_sc = (double *)qbx_pinned(64 * sizeof(double));       // shared pinned scratch: fetched at the point of use
    if (!h_sc) { qbx_set_error("pinned scratch allocation failed"); return QBX_ERR_NOMEM; }
    QBX_CUDA(cudaMemcpyAsync(h_sc, s->d_sc, 24 * sizeof(double), cudaMemcpyDeviceToHost, st));
    QBX_CUDA(cudaEventRecord(s->ev[3], st));
    QBX_CUDA(cudaStreamSynchronize(st));
    {
        float a = 0, f = 0, t = 0;
        cudaEventElapsedTime(&a, s->ev[0], s->ev[1]); cudaEventElapsedTime(&f, s->ev[1], s->ev[2]); cudaEventElapsedTime(&t, s->ev[0], s->ev[3]);
        if (damp > 0.0) {                                  // the first Fock build of a :DD step lies inside the solve bracket
            float f1 = 0;
            cudaEventElapsedTime(&f1, s->ev[4], s->ev[5]);
            a -= f1; f += f1;
        }
        s->t_solve += a * 1e-3; s->t_fock += f * 1e-3; s->t_total += t * 1e-3; s->nsteps += 1;
    }
    double resid = 0;
    for (int k = 0; k < nspin; ++k) {
        out[k] = 0.5 * h_sc[4 * k];
        resid += sqrt(h_sc[8 + 4 * k + 1] / (double)n2);
    }
    if (nspin == 1) out[1] = 0.0;
    out[2] = resid / nspin;
    out[3] = (nspin == 1 ? 2.0 : 1.0) * sqrt(h_sc[17] / (double)n2);
    return QBX_OK;
}

// copy the current (D, F, X^T R X) of every spin sector into history slot `slot`
extern "C" int qbx_scf_hist_store(qbx_scf *s, int slot)
{
    if (!s || slot < 0 || slot >= s->cap) { qbx_set_error("qbx_scf_hist_store: slot out of range"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    const size_t B = 2 * (size_t)s->n * s->n * sizeof(double);
    cudaStream_t st = qbx_stream();
    QBX_CUDA(cudaMemcpyAsync((char *)s->hD + slot * B, s->D, B, cudaMemcpyDeviceToDevice, st));
    QBX_CUDA(cudaMemcpyAsync((char *)s->hF + slot * B, s->F, B, cudaMemcpyDeviceToDevice, st));
    QBX_CUDA(cudaMemcpyAsync((char *)s->hR + slot * B, s->R, B, cudaMemcpyDeviceToDevice, st));
    return QBX_OK;
}

/* Gram matrices of m history slots of one spin sector (row-major m x m, host): Gdf[i][j] = <D_i, F_j> (what EDIIS and
 * ADIIS are built from, HartreeFock.jl:1273-1296) and Gee[i][j] = <e_i, e_j>, e = X^T (F D S - S D F) X (DIIS :1301-1315). */
extern "C" int qbx_scf_hist_gram(qbx_scf *s, int spin, int m, const int *slots, double *Gdf, double *Gee)
{
    if (!s || !slots || !Gdf || !Gee || m < 1 || m > s->cap || spin < 0 || spin > 1) { qbx_set_error("qbx_scf_hist_gram: bad argument"); return QBX_ERR_ARG; }
    for (int i = 0; i < m; ++i) if (slots[i] < 0 || slots[i] >= s->cap) { qbx_set_error("qbx_scf_hist_gram: slot out of range"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    cudaStream_t st = qbx_stream();
    const size_t n2 = (size_t)s->n * s->n;
    double *d = nullptr;
    QBX_CUDA(qbx_dmalloc(&d, (size_t)4 * m * m * sizeof(double)));
    QBX_CUDA(cudaMemsetAsync(d, 0, (size_t)4 * m * m * sizeof(double), st));
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            const double *Di = s->hD + (2 * (size_t)slots[i] + spin) * n2, *Fj = s->hF + (2 * (size_t)slots[j] + spin) * n2;
            k_dot2<<<64, 256, 0, st>>>(n2, Di, Fj, d + 2 * (i * m + j));
            if (j >= i) {
                const double *Ri = s->hR + (2 * (size_t)slots[i] + spin) * n2, *Rj = s->hR + (2 * (size_t)slots[j] + spin) * n2;
                k_dot2<<<64, 256, 0, st>>>(n2, Ri, Rj, d + 2 * m * m + 2 * (i * m + j));
            }
        }
    std::vector<double> h((size_t)4 * m * m);
    QBX_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    QBX_CUDA(cudaStreamSynchronize(st));
    qbx_pool_free(d);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            Gdf[i * m + j] = h[2 * (i * m + j)];
            Gee[i * m + j] = h[2 * m * m + 2 * (j >= i ? i * m + j : j * m + i)];
        }
    return QBX_OK;
}

// Fin[spin] = sum_i coef[i] F_hist[slots[i]]  (the extrapolated Fock matrix of xDIIS, :1318-1394)
extern "C" int qbx_scf_combine(qbx_scf *s, int spin, int m, const int *slots, const double *coef)
{
    if (!s || !slots || !coef || m < 1 || m > s->cap || spin < 0 || spin > 1) { qbx_set_error("qbx_scf_combine: bad argument"); return QBX_ERR_ARG; }
    for (int i = 0; i < m; ++i) if (slots[i] < 0 || slots[i] >= s->cap) { qbx_set_error("qbx_scf_combine: slot out of range"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    cudaStream_t st = qbx_stream();
    const size_t n2 = (size_t)s->n * s->n;
    QBX_CUDA(cudaMemsetAsync(s->Fin + spin * n2, 0, n2 * sizeof(double), st));
    for (int i = 0; i < m; ++i)
        k_accum<<<grid1(n2), 256, 0, st>>>(n2, coef[i], s->hF + (2 * (size_t)slots[i] + spin) * n2, s->Fin + spin * n2);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

// ------------------------------------------------------------------ f4: changeOrbitalBasis
// dense (ab|cd) on the device: the mode-2 store if there is one, else a temporary fill
static int dense_tensor(qbx_basis *b, double **T, bool *owned, cudaStream_t st)
{
    const int64_t N = b->nbf;
    if (b->mode == 2 && b->d_dense) { *T = b->d_dense; *owned = false; return QBX_OK; }
    if (qbx_dmalloc(T, (size_t)N * N * N * N * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        qbx_set_error("qbx_mo_transform: the dense N^4 tensor does not fit in device memory");
        return QBX_ERR_NOMEM;
    }
    *owned = true;
    int rc = b->eng ? b->eng->fill_tensor(*T, st, b->stats) : qbx_launch_generic_tensor(b->flat, *T, st);
    if (rc) { qbx_pool_free(*T); *T = nullptr; }
    return rc;
}

/* out[i,j,k,l] = sum_abcd (ab|cd) C[a,i] C[b,j] C[c,k] C[d,l]   (changeOrbitalBasis, Interface.jl:381-382)
 * C: nbf x nmo column-major (host); out: nmo^4 doubles, column-major (host). */
extern "C" int qbx_mo_transform(qbx_basis *b, int64_t nmo, const double *C, double *out, int64_t out_bytes)
{
    if (!b || !C || !out || nmo < 1) { qbx_set_error("qbx_mo_transform: bad argument"); return QBX_ERR_ARG; }
    const int64_t N = b->nbf, M = nmo;
    if (out_bytes < M * M * M * M * (int64_t)sizeof(double)) { qbx_set_error("qbx_mo_transform: output buffer smaller than nmo^4 * 8 bytes"); return QBX_ERR_ARG; }
    if (N > 2000 || M > 2000) { qbx_set_error("qbx_mo_transform: dimension too large"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    cudaStream_t st = qbx_stream();
    double *T = nullptr, *dC = nullptr, *A = nullptr, *B = nullptr;
    bool owned = false;
    if ((rc = dense_tensor(b, &T, &owned, st))) return rc;
    const size_t big = (size_t)std::max(N, M) * std::max(N, M) * std::max(N, M) * M;
    do {
        rc = QBX_ERR_NOMEM;
        if (qbx_dmalloc(&dC, (size_t)N * M * sizeof(double)) != cudaSuccess || qbx_dmalloc(&A, big * sizeof(double)) != cudaSuccess ||
            qbx_dmalloc(&B, big * sizeof(double)) != cudaSuccess) { cudaGetLastError(); qbx_set_error("qbx_mo_transform: out of device memory"); break; }
        cudaMemcpyAsync(dC, C, (size_t)N * M * sizeof(double), cudaMemcpyHostToDevice, st);
        // 1: A[i,(b c d)] = sum_a C[a,i] T[a,(b c d)]
        if ((rc = qbx_gemm(1, 0, (int)M, (int)(N * N * N), (int)N, 1.0, dC, (int)N, T, (int)N, 0.0, A, (int)M, st))) break;
        // 2: B[i,j,(c d)] = sum_b A[i,b,(c d)] C[b,j]          N^2 slabs of (M x N) . (N x M)
        if ((rc = qbx_gemm_batched(0, 0, (int)M, (int)M, (int)N, 1.0, A, (int)M, M * N, dC, (int)N, 0, 0.0, B, (int)M, M * M, (int)(N * N), st))) break;
        // 3: A[(i j),k,d] = sum_c B[(i j),c,d] C[c,k]           N slabs of (M^2 x N) . (N x M)
        if ((rc = qbx_gemm_batched(0, 0, (int)(M * M), (int)M, (int)N, 1.0, B, (int)(M * M), M * M * N, dC, (int)N, 0, 0.0, A, (int)(M * M), M * M * M, (int)N, st))) break;
        // 4: B[(i j k),l] = sum_d A[(i j k),d] C[d,l]
        if ((rc = qbx_gemm(0, 0, (int)(M * M * M), (int)M, (int)N, 1.0, A, (int)(M * M * M), dC, (int)N, 0.0, B, (int)(M * M * M), st))) break;
        if (cudaMemcpyAsync(out, B, (size_t)M * M * M * M * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { qbx_set_error(std::string("qbx_mo_transform: ") + cudaGetErrorString(cudaGetLastError())); rc = QBX_ERR_CUDA; break; }
        rc = QBX_OK;
    } while (0);
    qbx_pool_free(dC); qbx_pool_free(A); qbx_pool_free(B);
    if (owned) qbx_pool_free(T);
    return rc;
}

/* J[m,n] = sum_abcd (ab|cd) C1[a,m] C1[b,m] C2[c,n] C2[d,n]   (getJab, Interface.jl:384-393: the third element of
 * changeOrbitalBasis(twoBodyInt, C1, C2)).  out: nmo1 x nmo2 column-major (host). */
extern "C" int qbx_mo_coulomb_ab(qbx_basis *b, int64_t nmo1, const double *C1, int64_t nmo2, const double *C2, double *out)
{
    if (!b || !C1 || !C2 || !out || nmo1 < 1 || nmo2 < 1) { qbx_set_error("qbx_mo_coulomb_ab: bad argument"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    cudaStream_t st = qbx_stream();
    const int64_t N = b->nbf, N2 = N * N;
    double *T = nullptr, *P = nullptr, *Q = nullptr, *Y = nullptr, *J = nullptr;
    bool owned = false;
    if ((rc = dense_tensor(b, &T, &owned, st))) return rc;
    // P[(a b), m] = C1[a,m] C1[b,m], Q likewise: built on the host (N^2 nmo doubles), the contraction is two GEMMs
    std::vector<double> hP((size_t)N2 * nmo1), hQ((size_t)N2 * nmo2);
    for (int64_t m = 0; m < nmo1; ++m) for (int64_t bb = 0; bb < N; ++bb) for (int64_t a = 0; a < N; ++a) hP[m * N2 + bb * N + a] = C1[m * N + a] * C1[m * N + bb];
    for (int64_t m = 0; m < nmo2; ++m) for (int64_t bb = 0; bb < N; ++bb) for (int64_t a = 0; a < N; ++a) hQ[m * N2 + bb * N + a] = C2[m * N + a] * C2[m * N + bb];
    do {
        rc = QBX_ERR_NOMEM;
        if (qbx_dmalloc(&P, hP.size() * 8) != cudaSuccess || qbx_dmalloc(&Q, hQ.size() * 8) != cudaSuccess ||
            qbx_dmalloc(&Y, (size_t)N2 * nmo2 * 8) != cudaSuccess || qbx_dmalloc(&J, (size_t)nmo1 * nmo2 * 8) != cudaSuccess) {
            cudaGetLastError(); qbx_set_error("qbx_mo_coulomb_ab: out of device memory"); break;
        }
        cudaMemcpyAsync(P, hP.data(), hP.size() * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(Q, hQ.data(), hQ.size() * 8, cudaMemcpyHostToDevice, st);
        if ((rc = qbx_gemm(0, 0, (int)N2, (int)nmo2, (int)N2, 1.0, T, (int)N2, Q, (int)N2, 0.0, Y, (int)N2, st))) break;     // Y = T Q
        if ((rc = qbx_gemm(1, 0, (int)nmo1, (int)nmo2, (int)N2, 1.0, P, (int)N2, Y, (int)N2, 0.0, J, (int)nmo1, st))) break; // J = P^T Y
        if (cudaMemcpyAsync(out, J, (size_t)nmo1 * nmo2 * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            qbx_set_error(std::string("qbx_mo_coulomb_ab: ") + cudaGetErrorString(cudaGetLastError())); rc = QBX_ERR_CUDA; break;
        }
        rc = QBX_OK;
    } while (0);
    qbx_pool_free(P); qbx_pool_free(Q); qbx_pool_free(Y); qbx_pool_free(J);
    if (owned) qbx_pool_free(T);
    return rc;
}
