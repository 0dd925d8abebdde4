// The opaque handle of include/qbx.h and the few internals api.cu shares with scf.cu.
#pragma once
#include <memory>
#include <mutex>
#include <vector>

#include "engine.h"
#include "qbx_internal.h"

struct qbx_basis {
    std::mutex mu;
    // host copy of the boundary arrays
    int64_t nprim = 0, nbf = 0;
    std::vector<double> cen, xpn, bf_w;
    std::vector<int32_t> ang;
    std::vector<int64_t> bf_off, bf_prim;
    DevFlat flat{};                     // device copy (generic kernels)
    std::unique_ptr<Engine> eng;        // shell/class machinery (null if the basis is irregular)
    int mode = -1;                      // qbx_eri_store mode, -1 = nothing stored
    double *d_dense = nullptr;          // mode 2
    double *d_DJ = nullptr, *d_DK = nullptr, *d_G = nullptr;   // staging for host-pointer Fock builds
    int staged_nmat = 0;
    int nranks = 1;                     // shards the stored representation was cut into (qbx_eri_store)
    double stats[16] = {0};
};

// comm.cu
int qbx_comm_rank();
int qbx_comm_size();
int qbx_comm_allreduce(double *d_buf, size_t count, cudaStream_t s);


int qbx_ensure_init();                       // binds the device of this process (Julia tasks migrate between OS threads)
// getGcore on device pointers in the caller's numbering, incl. the all-reduce over the communicator's ranks
int qbx_fock_device(qbx_basis *b, int nmat, const double *dDJ, const double *dDK, double *dG, cudaStream_t s);
