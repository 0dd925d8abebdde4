/*
 * qbx.h -- C ABI of the B200-native ERI + Fock-build engine behind Quiqbox.jl.
 *
 * The reference (frankwswang/Quiqbox.jl v0.6.3) is pure Julia and has no FFI of its own;
 * the seams this library plugs into are two Julia methods:
 *
 *   seam 1  getOrbVectorIntegralCore!(::TwoBodyOrbIntegralInfo{T,3,T,<:CoulombInteractionSampler}, ptrVector)
 *           src/Integration/Framework.jl:640-698  (+ getOrbLayoutIntegralCore! :526-554)
 *   seam 2  getGcore(HeeI::AbstractArray{T,4}, DJ, DK)          src/HartreeFock.jl:305-319
 *
 * Everything is extern "C", plain pointers and sizes.  Every function returns 0 on success
 * and a non-zero code on failure, with text available from qbx_last_error(); no exception
 * crosses the boundary and the library never calls back into the host language.  Host
 * arrays belong to the caller and are only read/written during the call.  All device state
 * lives behind the opaque qbx_basis handle.  One process drives one GPU (qbx_init selects
 * it); multi-GPU runs are one process per GPU, each holding the shard (rank, nranks) of the
 * shell-quartet list, and sum their partial G matrices with an all-reduce (NCCL).
 *
 * Array conventions are the reference's: column-major, chemists' notation,
 * tensor[i,j,k,l] = (ij|kl), 0-based indices at this boundary.
 */
#ifndef QBX_H
#define QBX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qbx_basis qbx_basis;

/* libqbx.so is built with -fvisibility=hidden: the functions declared here are its ENTIRE dynamic symbol table (no C++
 * internals, no CUDA device stubs), so it can sit in one process next to other CUDA libraries (Julia + CUDA.jl, torch). */
#if defined(__GNUC__)
#define QBX_API __attribute__((visibility("default")))
#else
#define QBX_API
#endif

/* error codes */
enum { QBX_OK = 0, QBX_ERR_ARG = 1, QBX_ERR_CUDA = 2, QBX_ERR_STATE = 3, QBX_ERR_RANGE = 4, QBX_ERR_NOMEM = 5 };

/* Select and initialise CUDA device `device` for this process (idempotent).  n_dev_out (may
 * be NULL) receives the number of visible devices.  No reference counterpart (the reference
 * has no device); called once from the Julia glue's __init__. */
QBX_API int qbx_init(int device, int *n_dev_out);
QBX_API int qbx_shutdown(void);
QBX_API const char *qbx_last_error(void);

/* Basis ingestion = what MultiOrbitalData holds (src/OrbitalBases.jl:439-468) after
 * getOrbCorePointers/buildOrbCoreWeight! folded normalisation into the weights
 * (src/Integration/Framework.jl:748-817): a de-duplicated table of primitive Cartesian
 * Gaussians x^i y^j z^k exp(-xpn r^2) (prepareOrbitalInfoCore,
 * src/Integration/Engines/GaussianOrbitals.jl:65-78) and, per basis function, a CSR list of
 * (primitive index, final weight) (OrbCorePointer, Framework.jl:134-153).
 *   cen  3 x nprim column-major, xpn nprim, ang 3 x nprim,
 *   bf_off nbf+1 (0-based CSR), bf_prim / bf_w  bf_off[nbf] entries.
 * Shells are reconstructed inside (functions that share centre, exponents and radial
 * coefficients up to a per-component factor); functions that do not factor into shells
 * are kept as generic single functions. */
QBX_API int qbx_basis_create(int64_t nprim, const double *cen, const double *xpn, const int32_t *ang,
                     int64_t nbf, const int64_t *bf_off, const int64_t *bf_prim, const double *bf_w,
                     qbx_basis **out);
QBX_API int qbx_basis_destroy(qbx_basis *b);

/* info[0..15]: nbf, nshell, max l, class-path usable (1/0), n shell pairs, n unique shell
 * quartets in this rank's shard (after screening), n unique contracted ERI values in the
 * shard, stored bytes, n primitive quartets evaluated, rest reserved (0). */
QBX_API int qbx_basis_info(qbx_basis *b, int64_t *info);

/* seam 1, whole tensor: = elecRepulsions(bs) (src/Integration/Interface.jl:356-365 ->
 * getOrbVectorIntegralCore!, Framework.jl:640-698).  out: host, nbf^4 doubles, column-major,
 * all 8 permutational images written.  Fails without writing if out_bytes < nbf^4 * 8. */
QBX_API int qbx_eri_tensor(qbx_basis *b, double *out, int64_t out_bytes);

/* seam 1, single entries: = elecRepulsion(a,b,c,d) (Interface.jl:331-342 ->
 * getOrbLayoutIntegralCore!, Framework.jl:526-554).  ijkl: 4 x n, 0-based function indices;
 * any angular momentum (generic per-function kernel). */
QBX_API int qbx_eri_quartets(qbx_basis *b, int64_t n, const int64_t *ijkl, double *out);

/* Build this rank's device-resident ERI representation for repeated Fock builds: what
 * initializeHartreeFock obtains at src/HartreeFock.jl:189-191 and stores in
 * ElecHamiltonianConfig.twoBody (:143-151).
 *   screen_tol  Schwarz threshold on sqrt((ab|ab)(cd|cd)); 0 disables screening
 *   mode        0 = stored (packed unique ERIs kept in HBM), 1 = direct (recomputed in
 *               every qbx_fock_build), 2 = dense N^4 tensor (small N / irregular bases)
 *   rank,nranks shard of the cost-balanced shell-quartet list owned by this process
 * A basis with functions outside the s/p/d shell classes (l > 2, contractions over several centres) has no shell-quartet
 * lists: modes 0 and 1 are then served by the dense tensor when it is small (nbf^4 * 8 <= 16 GiB, one rank) and fail with
 * QBX_ERR_STATE and an explanatory message otherwise.
 * Thread safety: calls on ONE handle are serialised by a mutex inside the library; different handles may be used from
 * different host threads only one call at a time (the library has one stream and process-global kernel caches). */
QBX_API int qbx_eri_store(qbx_basis *b, double screen_tol, int mode, int rank, int nranks);

/* seam 2: = getGcore(HeeI, DJ, DK_m) for m = 0..nmat-1 sharing one DJ
 * (src/HartreeFock.jl:305-327; RHF nmat = 1 with DJ = 2D, DK = D; UHF nmat = 2 with
 * DJ = Da+Db, DK = Da, Db).  DJ: nbf^2, DK and G: nbf^2 * nmat, column-major, host.
 * G is Hermitian-filled.  DJ and DK must be symmetric (they are densities; the packed-store
 * digestion uses the 8-fold symmetry of the integrals together with D = D^T): a non-symmetric
 * argument is rejected with QBX_ERR_ARG in modes 0 and 1 (mode 2, the dense tensor, accepts any).
 * Multi-GPU (nranks > 1 in qbx_eri_store): when this process has joined a communicator of the same
 * size (qbx_comm_init) the partial matrices are summed INSIDE this call by one ncclAllReduce and
 * every rank receives the full G, as getGcore's callers expect (HartreeFock.jl:322-327); without a
 * communicator the result is this rank's PARTIAL G and the caller must sum over the ranks. */
QBX_API int qbx_fock_build(qbx_basis *b, int nmat, const double *DJ, const double *DK, double *G);

/* Communicator for nranks > 1: one process per GPU.  One rank obtains 128 bytes with
 * qbx_comm_unique_id and hands them to the others by whatever channel the host has (MPI.jl,
 * Distributed.jl, a file); then every rank calls qbx_comm_init(rank, nranks, id) after qbx_init.
 * NCCL is bound at run time (libnccl.so.2); nranks = 1 needs neither NCCL nor an id.
 * qbx_comm_info: rank and size of the communicator this process is in (0 and 1 if none). */
QBX_API int qbx_comm_unique_id(void *id128);
QBX_API int qbx_comm_init(int rank, int nranks, const void *id128);
QBX_API int qbx_comm_info(int *rank, int *nranks);
QBX_API int qbx_comm_destroy(void);

/* Same, device pointers and a caller stream (cudaStream_t as void*; NULL = the library's
 * stream); asynchronous with respect to the host.  The all-reduce over the communicator's ranks
 * (see qbx_fock_build) is enqueued on the same stream. */
QBX_API int qbx_fock_build_device(qbx_basis *b, int nmat, const double *dDJ, const double *dDK, double *dG,
                          void *stream);

/* Recompute this rank's shard of unique ERIs into the packed store (the ERI-throughput
 * step bench.py times).  Synchronous. */
QBX_API int qbx_eri_recompute(qbx_basis *b);

/* One-electron matrices needed to close an SCF (the first "next" row):
 * kind 0 overlap, 1 kinetic, 2 nuclear attraction (src/Integration/Interface.jl:46-312;
 * engines GaussianOrbitals.jl:94-363, 478-522).  Z: nnuc charges, R: 3 x nnuc. out: nbf^2. */
QBX_API int qbx_one_body(qbx_basis *b, int kind, int64_t nnuc, const double *Z, const double *R, double *out);

/* Boys function in isolation: out[(mmax+1)*t + m] = F_m(T[t])
 * (computeBoysSequence, src/Integration/Engines/BoysFunction.jl:67-77).
 * table != 0 uses the tabulated fast path of the class kernels (mmax <= 8). */
QBX_API int qbx_boys(int64_t n, const double *T, int mmax, int table, double *out);

/* Synthetic throughput sweep: nquartets contracted shell quartets of class
 * (la lb|lc ld) with uniform contraction degree K, generated from `seed` (centres uniform in
 * a 10-bohr cube, exponents log-uniform in [0.1,1e3], coefficients in [-1,1]).  secs: device
 * time of the ERI kernel, checksum: sum of all values, prim_quartets (may be NULL): primitive
 * shell quartets actually evaluated (primitive pairs whose prefactor underflows are dropped, so
 * this is <= nquartets * K^4).  sample_out (may be NULL): the first min(nsample, nquartets)
 * quartets' values and sample_geom their inputs, for oracle checks. */
QBX_API int qbx_prim_batch(int la, int lb, int lc, int ld, int K, int64_t nquartets, uint64_t seed,
                   double *secs, double *checksum, double *prim_quartets, int64_t nsample,
                   double *sample_out, double *sample_geom);

/* Asynchronous variant of qbx_eri_recompute: only enqueues the class kernels on the
 * library's stream (see qbx_set_stream). */
QBX_API int qbx_eri_recompute_async(qbx_basis *b);

/* Make every later launch of this process use `stream` (cudaStream_t as void*; NULL restores
 * the library's own stream).  Lets a host framework (torch.distributed + NCCL) order its
 * collectives and CUDA events with the library's kernels on one stream. */
QBX_API int qbx_set_stream(void *stream);

/* Per-class device times: runs one extra recompute with the class kernels serialised on the
 * library's stream and CUDA events around each of them (the normal recompute overlaps the
 * classes on side streams), then synchronises.
 * out[21][6]: la*1000+lb*100+lc*10+ld, seconds, shell quartets, primitive quartets,
 * model flops, component values.  Rows follow the canonical class order. */
QBX_API int qbx_class_stats(qbx_basis *b, double *out);

/* Measured FP64 FMA peak of the bound device (register-resident DFMA chains, all SMs):
 * the roofline denominator of the ERI kernels (MEASURED_PEAKS.json has no FP64 entry). */
QBX_API int qbx_fp64_peak(double *tflops);

/* Device allocations of the library (basis tables, task lists, the packed ERI store) come from
 * a size-keyed pool, so that the create -> store -> destroy cycle of a geometry scan or an
 * optimisation (reference: src/HartreeFock.jl:583-606 runs once per geometry) does not pay
 * cudaMalloc/cudaFree of a multi-GB store every time.  Environment QBX_POOL_GB = cap on idle
 * cached bytes (default 64, 0 = no pooling).  qbx_pool_trim returns the idle blocks to the
 * driver; counts (nullable) receives [0] reuses, [1] driver allocations, [2] idle bytes before
 * the trim. */
QBX_API int qbx_pool_trim(int64_t *counts);

/* counters since the last reset: [0] kernels launched, [1] device seconds in ERI kernels,
 * [2] device seconds in digestion kernels, [3] primitive quartets evaluated,
 * [4] model flops (SURVEY.md 8d counting rule), [5] bytes streamed by digestion, [6] stored-mode Fock builds that ran as
 * one CUDA-graph launch (a build is captured the second time it sees the same device buffers), [7..15] reserved */
QBX_API int qbx_stats(qbx_basis *b, double *out, int reset);

/* ---- the SCF step on the device.  getCDFE (src/HartreeFock.jl:392-403) -- solveFockMatrix
 * through X = S^(-1/2) (:39-56; `eigen` via cuSOLVER), getD (:296-302), getG / getF (:322-335; the Fock build is
 * qbx_fock_build_device incl. the all-reduce), getE (:339-350) -- plus the residual F D S - S D F (:1298) and the Gram
 * matrices the DIIS family is built from (:1273-1316), with every N x N matrix resident in HBM; the host keeps the
 * m x m coefficient problem and the stage logic.  S, Hcore: nbf^2 column-major (host); `history` = number of
 * (D, F, residual) slots kept on the device.  A store must exist on the basis (qbx_eri_store). */
typedef struct qbx_scf qbx_scf;
QBX_API int qbx_scf_create(qbx_basis *b, const double *S, const double *Hcore, int history, qbx_scf **out);
QBX_API int qbx_scf_destroy(qbx_scf *s);
/* which: 0 = the matrix the next step diagonalises (F_in), 1 = coefficient matrix C; nbf^2 doubles, host */
QBX_API int qbx_scf_set(qbx_scf *s, int which, int spin, const double *M);
/* which: 0 F_in, 1 C, 2 D, 3 F, 4 orbital energies (nbf doubles), 5 X, 6 = four doubles: device seconds since creation in
 * {eigen + transforms + density (a :DD step: incl. its first Fock build), the Fock builds, whole steps} and the step count */
QBX_API int qbx_scf_get(qbx_scf *s, int which, int spin, double *M);
/* One step for nspin sectors with nocc[spin] occupied orbitals.  from_coeff = 1 keeps the coefficients set with
 * qbx_scf_set instead of diagonalising F_in; damp > 0 is the damped step :DD (:1245-1270; two Fock builds).
 * out[0..1] = E per spin sector (getE), out[2] = mean over the sectors of RMS(F D S - S D F) (getErrorNrms, :1230),
 * out[3] = RMS change of the total density. */
QBX_API int qbx_scf_step(qbx_scf *s, int nspin, const int *nocc, int from_coeff, double damp, double *out);
QBX_API int qbx_scf_hist_store(qbx_scf *s, int slot);
/* Gdf[i][j] = <D_i, F_j>, Gee[i][j] = <e_i, e_j>, e = X^T (F D S - S D F) X over the m given slots (row-major m x m, host) */
QBX_API int qbx_scf_hist_gram(qbx_scf *s, int spin, int m, const int *slots, double *Gdf, double *Gee);
/* F_in[spin] = sum_i coef[i] F[slots[i]] */
QBX_API int qbx_scf_combine(qbx_scf *s, int spin, int m, const int *slots, const double *coef);

/* ---- changeOrbitalBasis (src/Integration/Interface.jl:376-407).
 * out[i,j,k,l] = sum (ab|cd) C[a,i] C[b,j] C[c,k] C[d,l], C: nbf x nmo column-major, out: nmo^4 doubles column-major
 * (host); four quarter transforms as FP64 GEMMs over the dense tensor on the device (the mode-2 store if present,
 * else built for the call: nbf^4 * 8 bytes must fit).  qbx_mo_coulomb_ab is the third element of the two-coefficient
 * method: J[m,n] = sum (ab|cd) C1[a,m] C1[b,m] C2[c,n] C2[d,n], nmo1 x nmo2 column-major. */
QBX_API int qbx_mo_transform(qbx_basis *b, int64_t nmo, const double *C, double *out, int64_t out_bytes);
QBX_API int qbx_mo_coulomb_ab(qbx_basis *b, int64_t nmo1, const double *C1, int64_t nmo2, const double *C2, double *out);

#ifdef __cplusplus
}
#endif
#endif
