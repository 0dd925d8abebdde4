// run-time bound cuBLAS / cuSOLVER wrappers (linalg.cu)
#pragma once
#include "qbx_internal.h"
int qbx_gemm(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double beta,
             double *C, int ldc, cudaStream_t s);
int qbx_gemm_batched(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, long long sa, const double *B, int ldb,
                     long long sb, double beta, double *C, int ldc, long long sc, int batch, cudaStream_t s);
int qbx_syevd(int n, double *A, double *w, cudaStream_t s);
