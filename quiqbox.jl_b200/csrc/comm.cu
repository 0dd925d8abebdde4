// The collective inside the boundary: one process per GPU, the partial Fock matrices of the
// ranks are summed by ONE ncclAllReduce(sum, f64, nmat N^2) at the end of qbx_fock_build / qbx_fock_build_device, on
// the library's stream, so that a host that only knows the reference's getGcore (src/HartreeFock.jl:305-327: the
// result IS the full G) needs no reducer of its own.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, re-using the copy a host framework such as torch has already
// loaded): libqbx.so has no link-time dependency on it and a single-GPU process never touches it.  Only the five
// entry points below are used; their C ABI has been stable since NCCL 2.0.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "qbx_internal.h"

namespace {
typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId_ { char internal[128]; };
enum { kNcclFloat64 = 8, kNcclSum = 0 };

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

std::mutex g_cmu;
Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_nranks = 1;

int bind_nccl()
{
    if (g_nccl.lib) return QBX_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { qbx_set_error(std::string("qbx_comm: cannot load libnccl.so.2: ") + dlerror()); return QBX_ERR_STATE; }
    Nccl n;
    n.lib = h;
    n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))dlsym(h, "ncclCommInitRank");
    n.AllReduce = (decltype(n.AllReduce))dlsym(h, "ncclAllReduce");
    n.CommDestroy = (decltype(n.CommDestroy))dlsym(h, "ncclCommDestroy");
    n.GetErrorString = (decltype(n.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.CommDestroy) {
        qbx_set_error("qbx_comm: libnccl.so.2 lacks a required entry point");
        return QBX_ERR_STATE;
    }
    g_nccl = n;
    return QBX_OK;
}

int nccl_fail(const char *what, int rc)
{
    qbx_set_error(std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"));
    return QBX_ERR_CUDA;
}
}   // namespace

int qbx_comm_rank() { return g_rank; }
int qbx_comm_size() { return g_nranks; }

// sum of `count` doubles over the ranks, in place, enqueued on `s`; a no-op for a single rank
int qbx_comm_allreduce(double *d_buf, size_t count, cudaStream_t s)
{
    if (g_nranks <= 1) return QBX_OK;
    if (!g_comm) { qbx_set_error("qbx_comm: communicator not initialised (qbx_comm_init)"); return QBX_ERR_STATE; }
    const int rc = g_nccl.AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, g_comm, s);
    return rc ? nccl_fail("ncclAllReduce", rc) : QBX_OK;
}

extern "C" int qbx_comm_unique_id(void *id128)
{
    if (!id128) { qbx_set_error("qbx_comm_unique_id: null argument"); return QBX_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_cmu);
    int rc = bind_nccl();
    if (rc) return rc;
    ncclUniqueId_ id;
    if ((rc = g_nccl.GetUniqueId(&id))) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return QBX_OK;
}

extern "C" int qbx_comm_init(int rank, int nranks, const void *id128)
{
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) { qbx_set_error("qbx_comm_init: bad argument"); return QBX_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_cmu);
    if (g_comm) { qbx_set_error("qbx_comm_init: a communicator already exists (qbx_comm_destroy first)"); return QBX_ERR_STATE; }
    if (nranks == 1) { g_rank = 0; g_nranks = 1; return QBX_OK; }
    int rc = bind_nccl();
    if (rc) return rc;
    ncclUniqueId_ id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    if ((rc = g_nccl.CommInitRank(&c, nranks, id, rank))) return nccl_fail("ncclCommInitRank", rc);
    g_comm = c; g_rank = rank; g_nranks = nranks;
    return QBX_OK;
}

extern "C" int qbx_comm_info(int *rank, int *nranks)
{
    if (rank) *rank = g_rank;
    if (nranks) *nranks = g_nranks;
    return QBX_OK;
}

extern "C" int qbx_comm_destroy(void)
{
    std::lock_guard<std::mutex> lk(g_cmu);
    if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
    g_rank = 0; g_nranks = 1;
    return QBX_OK;
}
