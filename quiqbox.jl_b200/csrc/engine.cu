// Host orchestration of the shell-quartet path (see engine.h).
#include <chrono>
#include <cstdio>
#include "engine.h"
#include "digest.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <atomic>
#include <memory>
#include <thread>
#include <numeric>
#include <random>
#include <stdlib.h>
#include <thrust/execution_policy.h>
#include <thrust/functional.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/binary_search.h>

// ------------------------------------------------------------------ class registry
#define QBX_DECL(a, b, c, d) extern const ClassOps qbx_ops_##a##b##c##d;
#define QBX_FOR_ALL_CLASSES(X)                                                                                  \
    X(0, 0, 0, 0)                                                                                               \
    X(1, 0, 0, 0) X(1, 0, 1, 0)                                                                                 \
    X(1, 1, 0, 0) X(1, 1, 1, 0) X(1, 1, 1, 1)                                                                   \
    X(2, 0, 0, 0) X(2, 0, 1, 0) X(2, 0, 1, 1) X(2, 0, 2, 0)                                                     \
    X(2, 1, 0, 0) X(2, 1, 1, 0) X(2, 1, 1, 1) X(2, 1, 2, 0) X(2, 1, 2, 1)                                       \
    X(2, 2, 0, 0) X(2, 2, 1, 0) X(2, 2, 1, 1) X(2, 2, 2, 0) X(2, 2, 2, 1) X(2, 2, 2, 2)
QBX_FOR_ALL_CLASSES(QBX_DECL)

static inline int pair_cls(int la, int lb) { return la * (la + 1) / 2 + lb; }

const ClassOps *qbx_class_ops(int bc, int kc)
{
    static const ClassOps *tab[QBX_NPAIRCLS][QBX_NPAIRCLS] = {{nullptr}};
    static bool init = false;
    if (!init) {
#define QBX_REG(a, b, c, d) tab[pair_cls(a, b)][pair_cls(c, d)] = &qbx_ops_##a##b##c##d;
        QBX_FOR_ALL_CLASSES(QBX_REG)
        init = true;
    }
    return (bc >= 0 && bc < QBX_NPAIRCLS && kc >= 0 && kc <= bc) ? tab[bc][kc] : nullptr;
}

static const int kClsLa[QBX_NPAIRCLS] = {0, 1, 1, 2, 2, 2};
static const int kClsLb[QBX_NPAIRCLS] = {0, 0, 1, 0, 1, 2};

// ------------------------------------------------------------------ flop model
static void comps_of(int l, int (*c)[3])
{
    int n = 0;
    for (int i = l; i >= 0; --i)
        for (int j = l - i; j >= 0; --j) { c[n][0] = i; c[n][1] = j; c[n][2] = l - i - j; ++n; }
}

double qbx_model_flops_prim(int la, int lb, int lc, int ld)
{
    const int L = la + lb + lc + ld, E = la + lb, F = lc + ld;
    double vrr = 0;
    int ce[45][3], cf[45][3];
    for (int e = 0; e <= E; ++e)
        for (int f = 0; f <= F; ++f) {
            if (e == 0 && f == 0) continue;
            comps_of(e, ce); comps_of(f, cf);
            for (int m = 0; m <= L - e - f; ++m)
                for (int a = 0; a < qbx_nc(e); ++a)
                    for (int b = 0; b < qbx_nc(f); ++b) {
                        int ax, low;
                        if (f > 0) { ax = cf[b][0] > 0 ? 0 : (cf[b][1] > 0 ? 1 : 2); low = cf[b][ax]; }
                        else { ax = ce[a][0] > 0 ? 0 : (ce[a][1] > 0 ? 1 : 2); low = ce[a][ax]; }
                        double c = 3;
                        if (low > 1) c += 4;
                        if (f > 0 && ce[a][ax] > 0) c += 2;
                        vrr += c;
                    }
        }
    double se = 0, sf = 0;
    for (int e = la; e <= E; ++e) se += qbx_nc(e);
    for (int f = lc; f <= F; ++f) sf += qbx_nc(f);
    return 84.0 + 25.0 + 3.0 * L + vrr + 2.0 * se * sf;
}

double qbx_model_flops_hrr(int la, int lb, int lc, int ld)
{
    const int E = la + lb, F = lc + ld;
    double sf = 0, h = 0;
    for (int f = lc; f <= F; ++f) sf += qbx_nc(f);
    for (int b = 1; b <= lb; ++b)
        for (int a = la; a <= E - b; ++a) h += 2.0 * qbx_nc(a) * qbx_nc(b) * sf;
    for (int d = 1; d <= ld; ++d)
        for (int c = lc; c <= F - d; ++c) h += 2.0 * qbx_nc(c) * qbx_nc(d) * qbx_nc(la) * qbx_nc(lb);
    return h;
}

// QBX_TRACE=2: stream synchronisation + elapsed time since the previous mark, on stderr (finds the slow enqueue-only step)
static void trace_mark(cudaStream_t s, const char *what, int a = -1, int b = -1)
{
    static const bool on = getenv("QBX_TRACE") && atoi(getenv("QBX_TRACE")) >= 2;
    if (!on) return;
    static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    const auto t0 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "    [qbx mark] %-22s %2d %2d  host %.3f ms  + device drain %.3f ms\n", what, a, b,
            1e3 * std::chrono::duration<double>(t0 - last).count(), 1e3 * std::chrono::duration<double>(t1 - t0).count());
    last = std::chrono::steady_clock::now();
}

// QBX_TRACE=1: host-side phase times on stderr
namespace {
struct TraceScope {
    const char *name;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit TraceScope(const char *n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("QBX_TRACE") && atoi(getenv("QBX_TRACE"))) {}
    ~TraceScope() { if (on) fprintf(stderr, "  [qbx trace] %-28s %.4f s\n", name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); }
};
}   // namespace

// ------------------------------------------------------------------ small kernels
namespace {

__global__ void k_diag_tasks(int n, int2 *t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = make_int2(i, i);
}

// Q[i] = sqrt(max_ab |(ab|ab)|) over the components of pair i
__global__ void k_schwarz(const double *vals, int n, int nab, double *Q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double m = 0.0;
    for (int ab = 0; ab < nab; ++ab) m = fmax(m, fabs(vals[(int64_t)(ab * nab + ab) * n + i]));
    Q[i] = sqrt(m);
}


// ---- primitive-pair records built on the device: the host uploads the shell table once (a few kB), a kernel counts
// the surviving primitive pairs of every shell pair, the host sorts the pairs by that count (the one small
// read-back), and a second kernel writes the AoS / SoA records in place.
struct DevShells { const double *cen; const int *xoff; const double *xpn, *coef; };

__device__ __forceinline__ double pair_prefactor(double pref, double a, double b, double ca, double cb, double ab2, double z)
{
    return pref * ca * cb * exp(-a * b / z * ab2) / z;        // same expression as the host path (build_pairset)
}

__global__ void k_pair_count(DevShells S, const int2 *sp, int np, double pref, int *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const int A = sp[i].x, B = sp[i].y;
    double ab2 = 0;
    for (int d = 0; d < 3; ++d) { const double t = S.cen[3 * A + d] - S.cen[3 * B + d]; ab2 += t * t; }
    int c = 0;
    for (int pa = S.xoff[A]; pa < S.xoff[A + 1]; ++pa)
        for (int pb = S.xoff[B]; pb < S.xoff[B + 1]; ++pb) {
            const double a = S.xpn[pa], b = S.xpn[pb];
            if (fabs(pair_prefactor(pref, a, b, S.coef[pa], S.coef[pb], ab2, a + b)) >= 1e-24) ++c;
        }
    cnt[i] = c;
}

__global__ void k_pair_fill(DevShells S, const int2 *sp, int np, double pref, const int *poff, const int2 *soa_idx, double *prim,
                            double *soa, double *geom)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= np) return;
    const int A = sp[n].x, B = sp[n].y;
    double ab2 = 0, ca[3], cb[3];
    for (int d = 0; d < 3; ++d) { ca[d] = S.cen[3 * A + d]; cb[d] = S.cen[3 * B + d]; const double t = ca[d] - cb[d]; ab2 += t * t; }
    for (int d = 0; d < 3; ++d) { geom[8 * (int64_t)n + d] = ca[d]; geom[8 * (int64_t)n + 3 + d] = ca[d] - cb[d]; }
    geom[8 * (int64_t)n + 6] = geom[8 * (int64_t)n + 7] = 0.0;
    const int64_t b0 = soa_idx[n].x, g = soa_idx[n].y;
    int c = 0;
    for (int pa = S.xoff[A]; pa < S.xoff[A + 1]; ++pa)
        for (int pb = S.xoff[B]; pb < S.xoff[B + 1]; ++pb) {
            const double a = S.xpn[pa], b = S.xpn[pb], z = a + b;
            const double K = pair_prefactor(pref, a, b, S.coef[pa], S.coef[pb], ab2, z);
            if (fabs(K) < 1e-24) continue;
            double v[8];
            v[0] = z;
            for (int d = 0; d < 3; ++d) v[1 + d] = (a * ca[d] + b * cb[d]) / z;
            v[4] = K; v[5] = b; v[6] = 0.5 / z; v[7] = 1.0 / z;
            double *r = prim + 8 * ((int64_t)poff[n] + c);
            for (int k = 0; k < 8; ++k) r[k] = v[k];
            for (int k = 0; k < QBX_SOA_NF; ++k) soa[b0 + ((int64_t)c * QBX_SOA_NF + k) * g] = v[k];
            ++c;
        }
}

// one warp per bra row
__global__ void k_count_tasks(const double *Qb, const double *Qk, int nb, int nk, int same, double tol, int *cnt)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (i >= nb) return;
    const int jmax = same ? i + 1 : nk;
    const double qi = Qb[i];
    int c = 0;
    for (int j = lane; j < jmax; j += 32) c += (qi * Qk[j] >= tol);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[i] = c;
}

// off[i] = cnt[0] + .. + cnt[i-1], off[n] = *total = the sum: one block walks over the array
__global__ void __launch_bounds__(1024) k_scan_counts(const int *cnt, int n, int64_t *off, int64_t *total)
{
    __shared__ int64_t wsum[32];
    __shared__ int64_t tile_sum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int64_t v = i < n ? cnt[i] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            const int64_t t = wsum[lane];
            int64_t z = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            wsum[lane] = z - t;
            if (lane == 31) tile_sum = z;
        }
        __syncthreads();
        if (i < n) off[i] = carry + wsum[w] + x - v;
        carry += tile_sum;
        __syncthreads();
    }
    if (threadIdx.x == 0) { off[n] = carry; if (total) *total = carry; }
}

// Sharding granule: rank r owns chunks r, r + nranks, ...  One chunk = one warp's worth of tasks.
// It must be much shorter than a bra row (up to ~6000 kets whose cost falls 81-fold along the
// row): with 4096-task chunks the round-robin aliased with the rows and rank 0 was 2x slower.
#define QBX_TASK_CHUNK 32
// one warp per bra row: compact the surviving kets in order; keep the chunks of this rank
__global__ void k_fill_tasks(const double *Qb, const double *Qk, int nb, int nk, int same, double tol,
                             const int64_t *rowoff, int rank, int nranks, int2 *tasks)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (i >= nb) return;
    const int jmax = same ? i + 1 : nk;
    const double qi = Qb[i];
    int64_t g = rowoff[i];
    // four windows of 32 kets per trip: the loads are issued together (a row is one warp walking serially over up to
    // ~8000 kets; with one load in flight per trip the 21 fill kernels of a store were pure latency)
    for (int j0 = 0; j0 < jmax; j0 += 128) {
        double q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int j = j0 + 32 * u + lane; q[u] = j < jmax ? Qk[j] : -1.0; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u + lane;
            const bool pass = j < jmax && qi * q[u] >= tol;
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const int64_t gg = g + __popc(mask & ((1u << lane) - 1u));
                const int64_t c = gg / QBX_TASK_CHUNK;
                if (c % nranks == rank) tasks[(c / nranks) * QBX_TASK_CHUNK + gg % QBX_TASK_CHUNK] = make_int2(i, j);
            }
            g += __popc(mask);
        }
    }
}

__global__ void k_task_cost(const int2 *tasks, int64_t n, const int *poffb, const int *poffk, double *sum)
{
    __shared__ double red[256];
    double acc = 0.0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int2 t = tasks[q];
        acc += (double)(poffb[t.x + 1] - poffb[t.x]) * (double)(poffk[t.y + 1] - poffk[t.y]);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(sum, red[0]);
}

// G = (Jt + Jt^T) - (Kt + Kt^T), written in the caller's (external) numbering.  *bad != 0: a density was not
// symmetric (k_permute_in) -- the half-accumulators mean nothing then, and G is poisoned with NaN rather than
// returned silently wrong (the host-pointer entry point rejects such input before it gets here).
__global__ void k_finish_G(int64_t Nint, int64_t Next, const int *ext_of_int, int nmat, const double *Jt, const double *Kt,
                           double *G, const int *bad)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= Nint * Nint) return;
    const int64_t i = e % Nint, j = e / Nint, et = j + Nint * i;
    const int ie = ext_of_int[i], je = ext_of_int[j];
    if (ie < 0 || je < 0) return;
    const double jv = *bad ? nan("") : Jt[e] + Jt[et];
    for (int m = 0; m < nmat; ++m)
        G[m * Next * Next + ie + Next * je] = jv - (Kt[m * Nint * Nint + e] + Kt[m * Nint * Nint + et]);
}

// external <-> internal function numbering (internal = complete Cartesian shells, shell by shell); flags a density
// that is not symmetric (relative 1e-10: densities C C^T are symmetric to rounding)
__global__ void k_permute_in(int64_t Next, int64_t Nint, const int *ext_of_int, const double *Dext, double *Dint, int *bad)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= Nint * Nint) return;
    const int i = ext_of_int[e % Nint], j = ext_of_int[e / Nint];
    double v = 0.0;
    if (i >= 0 && j >= 0) {
        v = Dext[i + Next * j];
        const double w = Dext[j + Next * i];
        if (!(fabs(v - w) <= 1e-10 * fmax(1.0, fmax(fabs(v), fabs(w))))) atomicOr(bad, 1);
    }
    Dint[e] = v;
}

// cost of every 32-task chunk = primitive quartets behind its tasks
__global__ void k_chunk_cost(const int2 *tasks, const int *gt_bra, const int *gt_grp, int64_t n, const int *poffb,
                             const int *poffk, float *cost, int *nheavy)
{
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // blockDim multiple of 32
    float c = 0.f;
    if (q < n) {
        int b, k;
        if (tasks) { b = tasks[q].x; k = tasks[q].y; } else { b = gt_bra[q]; k = gt_grp[q]; }
        if (k >= 0) c = (float)(poffb[b + 1] - poffb[b]) * (float)(poffk[k + 1] - poffk[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));   // a warp runs at its slowest lane
    if ((threadIdx.x & 31) == 0 && q < n + 31) {
        cost[q / 32] = c;
        if (nheavy && c > QBX_HEAVY_TASK) atomicAdd(nheavy, 1);
    }
}

__global__ void k_sum(const double *v, int64_t n, double *sum)
{
    __shared__ double red[256];
    double acc = 0.0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) acc += v[q];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(sum, red[0]);
}

}  // namespace

void qbx_scan_counts(const int *cnt, int n, int64_t *off, int64_t *d_total, cudaStream_t s)
{
    k_scan_counts<<<1, 1024, 0, s>>>(cnt, n, off, d_total);
}

namespace {
// temporary storage of the thrust algorithms comes from the library's pool (no cudaMalloc/cudaFree)
struct PoolAllocator {
    typedef char value_type;
    char *allocate(std::ptrdiff_t n)
    {
        void *p = nullptr;
        if (qbx_pool_malloc(&p, (size_t)n) != cudaSuccess) throw std::bad_alloc();
        return (char *)p;
    }
    void deallocate(char *p, size_t) { qbx_pool_free_async(p); }
};
}   // namespace

// Enqueue-only: nothing here waits for the device.  *d_nheavy (device, zeroed by the caller, may
// be null) receives the number of chunks above QBX_HEAVY_TASK = the leading entries of the order.
int qbx_chunk_order(const int2 *tasks, const int *gt_bra, const int *gt_grp, int64_t n, const int *poff_bra,
                    const int *poff_ket, int **order_out, int *d_nheavy, cudaStream_t s)
{
    *order_out = nullptr;
    const int64_t nchunk = (n + 31) / 32;
    if (nchunk <= 1) return QBX_OK;
    float *cost = nullptr;
    QBX_CUDA(qbx_dmalloc(&cost, nchunk * sizeof(float)));
    QBX_CUDA(qbx_dmalloc(order_out, nchunk * sizeof(int)));
    const int64_t threads = nchunk * 32;
    k_chunk_cost<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(tasks, gt_bra, gt_grp, n, poff_bra, poff_ket, cost, d_nheavy);
    QBX_CUDA(cudaGetLastError());
    try {
        PoolAllocator alloc;
        auto pol = thrust::cuda::par_nosync(alloc).on(s);
        thrust::sequence(pol, *order_out, *order_out + nchunk);
        thrust::stable_sort_by_key(pol, cost, cost + nchunk, *order_out, thrust::greater<float>());
    } catch (const std::exception &e) {
        qbx_set_error(std::string("chunk order: ") + e.what());
        qbx_pool_free(cost);
        return QBX_ERR_CUDA;
    }
    qbx_pool_free_async(cost);
    return QBX_OK;
}

// ------------------------------------------------------------------ shell reconstruction
static int comp_index(const int32_t *a) { return CIDX(a[1], a[2]); }

Engine *Engine::create(int64_t nprim, const double *cen, const double *xpn, const int32_t *ang, int64_t nbf,
                       const int64_t *bf_off, const int64_t *bf_prim, const double *bf_w)
{
    (void)nprim;
    TraceScope tr_all("create: total");
    std::unique_ptr<TraceScope> tr(new TraceScope("create: shell reconstruction"));
    std::vector<HostShell> shells;
    struct Key {
        double c[3]; int l; std::vector<double> x;
        bool operator<(const Key &o) const {
            if (l != o.l) return l < o.l;
            for (int d = 0; d < 3; ++d) if (c[d] != o.c[d]) return c[d] < o.c[d];
            return x < o.x;
        }
    };
    std::map<Key, std::vector<int>> open;
    for (int64_t f = 0; f < nbf; ++f) {
        const int64_t p0 = bf_off[f], p1 = bf_off[f + 1];
        const int64_t q0 = bf_prim[p0];
        Key k;
        for (int d = 0; d < 3; ++d) k.c[d] = cen[3 * q0 + d];
        k.l = ang[3 * q0] + ang[3 * q0 + 1] + ang[3 * q0 + 2];
        if (k.l > QBX_MAX_L) return nullptr;
        for (int64_t p = p0; p < p1; ++p) {
            const int64_t q = bf_prim[p];
            for (int d = 0; d < 3; ++d)
                if (cen[3 * q + d] != k.c[d] || ang[3 * q + d] != ang[3 * q0 + d]) return nullptr;   // not a shell function
            k.x.push_back(xpn[q]);
        }
        const int comp = comp_index(ang + 3 * q0);
        int found = -1;
        double ratio = 1.0;
        for (int si : open[k]) {
            HostShell &s = shells[si];
            if (s.bf[comp] >= 0) continue;
            bool ok = true, have = false;
            double r = 0.0;
            for (size_t p = 0; p < k.x.size() && ok; ++p) {
                const double w = bf_w[p0 + p], c = s.coef[p];
                if (c == 0.0) { ok = (w == 0.0); continue; }
                const double rp = w / c;
                if (!have) { r = rp; have = true; }
                else ok = fabs(rp - r) <= 1e-12 * fabs(r);
            }
            if (ok && have) { found = si; ratio = r; break; }
        }
        if (found < 0) {
            HostShell s;
            s.l = k.l;
            for (int d = 0; d < 3; ++d) s.cen[d] = k.c[d];
            s.xpn = k.x;
            s.coef.assign(bf_w + p0, bf_w + p1);
            shells.push_back(s);
            found = (int)shells.size() - 1;
            open[k].push_back(found);
            ratio = 1.0;
        }
        shells[found].bf[comp] = (int)f;
        shells[found].scale[comp] = ratio;
    }
    // internal order: by l, then longer contractions first (so that, for a fixed first shell, the
    // pairs (C,D) listed with D ascending have non-increasing primitive counts), then input order
    std::stable_sort(shells.begin(), shells.end(), [](const HostShell &a, const HostShell &b) {
        return a.l != b.l ? a.l < b.l : a.xpn.size() > b.xpn.size();
    });
    tr.reset();
    return from_shells(shells, nbf, false);
}

Engine *Engine::from_shells(const std::vector<HostShell> &shells, int64_t nbf, bool pair_adjacent)
{
    Engine *e = new Engine;
    e->shells_ = shells;
    e->nbf_ = nbf;
    for (auto &s : shells) e->maxl_ = std::max(e->maxl_, s.l);
    if (e->upload(pair_adjacent) != QBX_OK) { delete e; return nullptr; }
    return e;
}

// Primitive-pair records of the pair classes, computed on the device (k_pair_count / k_pair_fill) in two enqueue-only
// phases, so that a qbx_basis_create waits for the device twice and not twice per class: `count` uploads the shell
// pairs and counts their primitive pairs; after the ONE read-back the host sorts the pairs of every class by that
// count and `fill` writes the AoS and SoA records in place.  [The host-thread variant of round 1 -- 6 of the 7 ms of
// qbx_basis_create -- was deleted after the A/B of round 2: host-buffer step 52.1 -> 49.1 ms,
// profiles/r02/probe_ab_head_of_round1.log.]
struct PairBuild {
    std::vector<int2> sp0, shells, soa_idx;     // host copies stay alive until the stream has been synchronised
    std::vector<int> cnt, poff;
    int2 *d_sp = nullptr;
    int *d_cnt = nullptr;
};

static int pairset_count(const DevShells &S, const std::vector<std::pair<int, int>> &sp, PairBuild &w, cudaStream_t s)
{
    const size_t np_ = sp.size();
    const double pref = sqrt(2.0) * pow(M_PI, 1.25);
    w.sp0.resize(np_);
    for (size_t i = 0; i < np_; ++i) w.sp0[i] = make_int2(sp[i].first, sp[i].second);
    w.cnt.assign(np_, 0);
    QBX_CUDA(qbx_dmalloc(&w.d_sp, std::max<size_t>(1, np_) * sizeof(int2)));
    QBX_CUDA(qbx_dmalloc(&w.d_cnt, std::max<size_t>(1, np_) * sizeof(int)));
    if (np_) {
        QBX_CUDA(cudaMemcpyAsync(w.d_sp, w.sp0.data(), np_ * sizeof(int2), cudaMemcpyHostToDevice, s));
        k_pair_count<<<(unsigned)((np_ + 127) / 128), 128, 0, s>>>(S, w.d_sp, (int)np_, pref, w.d_cnt);
        QBX_CUDA(cudaMemcpyAsync(w.cnt.data(), w.d_cnt, np_ * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    return QBX_OK;
}

static int pairset_fill(const DevShells &S, int la, int lb, bool sort_by_nprim, PairBuild &w, DevPairSet &out, cudaStream_t s)
{
    const size_t np_ = w.sp0.size();
    out.la = la; out.lb = lb; out.npair = (int)np_; out.nprim = 0;
    const double pref = sqrt(2.0) * pow(M_PI, 1.25);
    const std::vector<int> &cnt = w.cnt;
    std::vector<int> order(np_);
    if (sort_by_nprim) {
        // stable, descending count: a counting sort (the counts are at most K_a * K_b, a few hundred)
        int cmax = 0;
        for (int c : cnt) cmax = std::max(cmax, c);
        std::vector<int> start(cmax + 2, 0);
        for (int c : cnt) ++start[cmax - c + 1];
        for (int c = 0; c <= cmax; ++c) start[c + 1] += start[c];
        for (size_t i = 0; i < np_; ++i) order[start[cmax - cnt[i]]++] = (int)i;
    } else {
        std::iota(order.begin(), order.end(), 0);
    }
    w.shells.assign(np_, make_int2(0, 0));
    w.poff.assign(np_ + 1, 0);
    out.h_nprim.resize(np_);
    for (size_t n = 0; n < np_; ++n) {
        w.shells[n] = w.sp0[order[n]];
        out.h_nprim[n] = cnt[order[n]];
        w.poff[n + 1] = w.poff[n] + cnt[order[n]];
    }
    out.nprim = w.poff[np_];
    const size_t n_prim = 8 * (size_t)w.poff[np_], n_soa = (size_t)QBX_SOA_NF * w.poff[np_];
    w.soa_idx.resize(np_);
    {
        size_t base = 0, g0 = 0;
        while (g0 < np_) {
            size_t g1 = g0;
            while (g1 < np_ && out.h_nprim[g1] == out.h_nprim[g0]) ++g1;
            const size_t g = g1 - g0;
            for (size_t j = g0; j < g1; ++j) w.soa_idx[j] = make_int2((int)(base + (j - g0)), (int)g);
            base += g * (size_t)out.h_nprim[g0] * QBX_SOA_NF;
            g0 = g1;
        }
    }
    QBX_CUDA(qbx_dmalloc(&out.soa, std::max<size_t>(1, n_soa) * sizeof(double)));
    QBX_CUDA(qbx_dmalloc(&out.soa_idx, std::max<size_t>(1, np_) * sizeof(int2)));
    QBX_CUDA(qbx_dmalloc(&out.shells, std::max<size_t>(1, np_) * sizeof(int2)));
    QBX_CUDA(qbx_dmalloc(&out.prim_off, w.poff.size() * sizeof(int)));
    QBX_CUDA(qbx_dmalloc(&out.geom, std::max<size_t>(1, 8 * np_) * sizeof(double)));
    QBX_CUDA(qbx_dmalloc(&out.prim, std::max<size_t>(1, n_prim) * sizeof(double)));
    QBX_CUDA(qbx_dmalloc(&out.schwarz, std::max<size_t>(1, np_) * sizeof(double)));
    QBX_CUDA(cudaMemcpyAsync(out.prim_off, w.poff.data(), w.poff.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    if (np_) {
        QBX_CUDA(cudaMemcpyAsync(out.soa_idx, w.soa_idx.data(), np_ * sizeof(int2), cudaMemcpyHostToDevice, s));
        QBX_CUDA(cudaMemcpyAsync(out.shells, w.shells.data(), np_ * sizeof(int2), cudaMemcpyHostToDevice, s));
        k_pair_fill<<<(unsigned)((np_ + 127) / 128), 128, 0, s>>>(S, out.shells, (int)np_, pref, out.prim_off, out.soa_idx, out.prim,
                                                                  out.soa, out.geom);
        QBX_CUDA(cudaGetLastError());
    }
    qbx_pool_free_async(w.d_sp); qbx_pool_free_async(w.d_cnt);
    w.d_sp = nullptr; w.d_cnt = nullptr;
    return QBX_OK;
}

int Engine::upload(bool pair_adjacent)
{
    // Everything below is enqueued on the library's stream from host vectors that live until the end of this function;
    // the host waits for the device three times: primitive-pair counts, group counts, end.
    cudaStream_t st = qbx_stream();
    const size_t ns = shells_.size();
    std::vector<int> bf(6 * ns);
    std::vector<double> sc(6 * ns);
    for (size_t s = 0; s < ns; ++s)
        for (int c = 0; c < 6; ++c) { bf[6 * s + c] = shells_[s].bf[c]; sc[6 * s + c] = shells_[s].scale[c]; }
    QBX_CUDA(qbx_dmalloc(&d_shell_bf_, std::max<size_t>(1, bf.size()) * sizeof(int)));
    QBX_CUDA(qbx_dmalloc(&d_shell_scale_, std::max<size_t>(1, sc.size()) * sizeof(double)));
    QBX_CUDA(cudaMemcpyAsync(d_shell_bf_, bf.data(), bf.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    QBX_CUDA(cudaMemcpyAsync(d_shell_scale_, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<int> first(ns), ext;
    for (size_t s = 0; s < ns; ++s) {
        first[s] = (int)ext.size();
        for (int c = 0; c < qbx_nc(shells_[s].l); ++c) ext.push_back(shells_[s].bf[c]);
    }
    nint_ = (int64_t)ext.size();
    QBX_CUDA(qbx_dmalloc(&d_shell_first_, std::max<size_t>(1, ns) * sizeof(int)));
    QBX_CUDA(qbx_dmalloc(&d_ext_of_int_, std::max<size_t>(1, ext.size()) * sizeof(int)));
    QBX_CUDA(cudaMemcpyAsync(d_shell_first_, first.data(), ns * sizeof(int), cudaMemcpyHostToDevice, st));
    QBX_CUDA(cudaMemcpyAsync(d_ext_of_int_, ext.data(), ext.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    std::vector<std::pair<int, int>> sp[QBX_NPAIRCLS];
    if (pair_adjacent) {
        for (size_t s = 0; s + 1 < ns; s += 2) {
            int a = (int)s, b = (int)s + 1;
            if (shells_[a].l < shells_[b].l) std::swap(a, b);
            sp[pair_cls(shells_[a].l, shells_[b].l)].push_back({a, b});
        }
    } else {
        // A major, B ascending (shells are sorted by l, so a >= b implies la >= lb): consecutive pairs share the
        // first shell, which the digestion's segmented warp sums rely on (digest.cuh).  [Listing the pairs diagonal
        // by diagonal, so that the 32 kets of a warp hold 32 different shells C and D, was measured in round 2
        // and lost: it needs 32 shells of one kind, (H2O)16 has 16 -- profiles/r02/digest_history.md.]
        for (auto &v : sp) v.reserve(ns * (ns + 1) / 2);
        for (size_t a = 0; a < ns; ++a)
            for (size_t b = 0; b <= a; ++b) sp[pair_cls(shells_[a].l, shells_[b].l)].push_back({(int)a, (int)b});
    }
    std::vector<int> first_h(ns);
    { int acc = 0; for (size_t s = 0; s < ns; ++s) { first_h[s] = acc; acc += qbx_nc(shells_[s].l); } }
    DevShells S{nullptr, nullptr, nullptr, nullptr};
    void *d_tab[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<double> cen(3 * ns), xp, cf;
    std::vector<int> xoff(ns + 1, 0);
    {
        for (size_t i = 0; i < ns; ++i) {
            for (int d = 0; d < 3; ++d) cen[3 * i + d] = shells_[i].cen[d];
            xp.insert(xp.end(), shells_[i].xpn.begin(), shells_[i].xpn.end());
            cf.insert(cf.end(), shells_[i].coef.begin(), shells_[i].coef.end());
            xoff[i + 1] = (int)xp.size();
        }
        const size_t bytes[4] = {cen.size() * sizeof(double), xoff.size() * sizeof(int), xp.size() * sizeof(double), cf.size() * sizeof(double)};
        const void *src[4] = {cen.data(), xoff.data(), xp.data(), cf.data()};
        for (int i = 0; i < 4; ++i) {
            QBX_CUDA(qbx_pool_malloc(&d_tab[i], std::max<size_t>(8, bytes[i])));
            if (bytes[i]) QBX_CUDA(cudaMemcpyAsync(d_tab[i], src[i], bytes[i], cudaMemcpyHostToDevice, st));
        }
        S = DevShells{(const double *)d_tab[0], (const int *)d_tab[1], (const double *)d_tab[2], (const double *)d_tab[3]};
    }
    PairBuild work[QBX_NPAIRCLS];
    std::vector<int4> info[QBX_NPAIRCLS];
    int rc;
    {
        TraceScope trc("create: pair counts");
        for (int pc = 0; pc < QBX_NPAIRCLS; ++pc)
            if ((rc = pairset_count(S, sp[pc], work[pc], st))) return rc;
        QBX_CUDA(cudaStreamSynchronize(st));                 // read-back 1: primitive pairs per shell pair, all classes
    }
    {
        TraceScope trf("create: pair fill (enqueue)");
        for (int pc = 0; pc < QBX_NPAIRCLS; ++pc) {
            // sorted by primitive count (stable): ties keep the (A major, B ascending) order, so inside an
            // equal-count group a run of pairs shares A and walks over consecutive B
            DevPairSet &P = pairs_[pc];
            if ((rc = pairset_fill(S, kClsLa[pc], kClsLb[pc], !pair_adjacent, work[pc], P, st))) return rc;
            const std::vector<int2> &sh = work[pc].shells;
            info[pc].resize(P.npair);
            for (int i = 0; i < P.npair; ++i) info[pc][i] = make_int4(sh[i].x, sh[i].y, first_h[sh[i].x], first_h[sh[i].y]);
            QBX_CUDA(qbx_dmalloc(&P.info, std::max<size_t>(1, info[pc].size()) * sizeof(int4)));
            if (P.npair) QBX_CUDA(cudaMemcpyAsync(P.info, info[pc].data(), info[pc].size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        }
    }
    // ket-side general-contraction sharing for the (xs|ss) classes (QBX_GC=0 switches it off)
    if (!pair_adjacent && pairs_[0].npair && !(getenv("QBX_GC") && atoi(getenv("QBX_GC")) == 0)) {
        TraceScope trg("create: group build");
        if ((rc = qbx_group_build(shells_, work[0].shells, groups_, pairs_[0].shells))) return rc;
        use_groups_ = groups_.ng > 0 && groups_.ng < pairs_[0].npair;     // only when something is shared
    }
    {
        TraceScope trs("create: final sync");
        QBX_CUDA(cudaStreamSynchronize(st));                 // the host vectors above go out of scope
    }
    for (void *p : d_tab) qbx_pool_free_async(p);
    QBX_CUDA(cudaEventCreate(&ev0_));
    QBX_CUDA(cudaEventCreate(&ev1_));
    return QBX_OK;
}

Engine::~Engine()
{
    QbxPoolFreeScope one_sync;
    release_store();
    qbx_group_free(groups_);
    for (auto &p : pairs_) { qbx_pool_free(p.shells); qbx_pool_free(p.prim_off); qbx_pool_free(p.geom); qbx_pool_free(p.prim); qbx_pool_free(p.schwarz); qbx_pool_free(p.soa); qbx_pool_free(p.soa_idx); qbx_pool_free(p.info); }
    qbx_pool_free(d_shell_bf_); qbx_pool_free(d_shell_scale_); qbx_pool_free(d_shell_first_); qbx_pool_free(d_ext_of_int_); qbx_pool_free(d_Dint_);
    qbx_pool_free(chunk_); qbx_pool_free(d_Jt_); qbx_pool_free(d_Kt_); qbx_pool_free(d_counters_); qbx_pool_free(d_bad_);

    for (auto &e : cls_ev_) if (e) cudaEventDestroy(e);
    for (int i = 0; i < kSide; ++i) { if (side_[i]) cudaStreamDestroy(side_[i]); if (side_ev_[i]) cudaEventDestroy(side_ev_[i]); }
    if (fork_ev_) cudaEventDestroy(fork_ev_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
}

void Engine::release_store()
{
    QbxPoolFreeScope one_sync;
    ++store_gen_;
    drop_fock_graph();
    for (int b = 0; b < QBX_NPAIRCLS; ++b)
        for (int k = 0; k < QBX_NPAIRCLS; ++k) {
            qbx_pool_free(tasks_[b][k].tasks); qbx_pool_free(tasks_[b][k].gt_bra); qbx_pool_free(tasks_[b][k].gt_grp); qbx_pool_free(tasks_[b][k].gt_off); qbx_pool_free(tasks_[b][k].order);
            tasks_[b][k] = TaskList();
            qbx_pool_free(vals_[b][k]); vals_[b][k] = nullptr;
        }
    mode_ = -1;
    n_quartets_ = n_values_ = stored_bytes_ = 0;
    n_primq_ = model_flops_ = 0;
}

void Engine::info(int64_t *info) const
{
    info[1] = (int64_t)shells_.size();
    info[2] = maxl_;
    info[3] = 1;
    int64_t np = 0;
    for (auto &p : pairs_) np += p.npair;
    info[4] = np;
    info[5] = n_quartets_;
    info[6] = n_values_;
    info[7] = stored_bytes_;
    info[8] = (int64_t)n_primq_;
    info[9] = (int64_t)model_flops_;
}

// ------------------------------------------------------------------ ERI launches
int Engine::eri_args(int bc, int kc, const int2 *tasks, int64_t n, double *out, cudaStream_t s, ClassArgs &a)
{
    a.bra = pairs_[bc].view(); a.ket = pairs_[kc].view();
    a.tasks = tasks; a.ntasks = n; a.out = out;
    a.shell_scale = d_shell_scale_;
    a.boys = qbx_boys_table();
    if (!d_counters_) QBX_CUDA(qbx_dmalloc(&d_counters_, 1024 * sizeof(unsigned int)));
    a.counter = d_counters_ + counter_next_;
    counter_next_ = (counter_next_ + 1) % 1024;
    QBX_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), s));
    a.order = nullptr;
    return QBX_OK;
}

int Engine::run_eri(int bc, int kc, const int2 *tasks, int64_t n, double *out, cudaStream_t s, const int *order)
{
    const ClassOps *ops = qbx_class_ops(bc, kc);
    if (!ops) { qbx_set_error("internal: no kernel for this class"); return QBX_ERR_STATE; }
    ClassArgs a;
    int rc0 = eri_args(bc, kc, tasks, n, out, s, a);
    if (rc0) return rc0;
    a.order = order;
    // Large classes (>= coop_min contracted accumulators per quartet) go to the warp-cooperative
    // kernel; QBX_COOP_MIN_ACC overrides the threshold (0 = every class, for tests).
    static const int coop_min = std::min(QBX_COOP_ACC, getenv("QBX_COOP_MIN_ACC") ? atoi(getenv("QBX_COOP_MIN_ACC")) : QBX_COOP_ACC);
    const int nacc = NCSUM(ops->la, ops->la + ops->lb) * NCSUM(ops->lc, ops->lc + ops->ld);
    if (nacc >= coop_min) {
        const int rc = qbx_launch_eri_coop(ops->la, ops->lb, ops->lc, ops->ld, a, s);
        if (rc >= 0) return rc;
    }
    return ops->eri(a, s);
}

int Engine::ensure_schwarz(cudaStream_t s)
{
    if (have_schwarz_) return QBX_OK;
    // the six diagonal classes (ab|ab) are independent: side streams, nothing waits on the host
    int rc = fork(s);
    if (rc) return rc;
    std::vector<void *> scratch;
    int k = 0;
    for (int pc = 0; pc < QBX_NPAIRCLS; ++pc) {
        DevPairSet &P = pairs_[pc];
        if (P.npair == 0) continue;
        const ClassOps *ops = qbx_class_ops(pc, pc);
        cudaStream_t cs = side_[k++ % kSide];
        int2 *t = nullptr; double *v = nullptr;
        QBX_CUDA(qbx_dmalloc(&t, P.npair * sizeof(int2)));
        QBX_CUDA(qbx_dmalloc(&v, (size_t)ops->ncomp * P.npair * sizeof(double)));
        scratch.push_back(t); scratch.push_back(v);
        k_diag_tasks<<<(P.npair + 127) / 128, 128, 0, cs>>>(P.npair, t);
        // a diagonal (ab|ab) of two long contractions is thousands of primitive quartets: one WARP per pair (one thread
        // per pair, the first version, cost the host-buffer step 2 ms: 54.3 vs 52.2 ms in the A/B of round 2)
        if (ops->eri_split) {
            ClassArgs a;
            if ((rc = eri_args(pc, pc, t, P.npair, v, cs, a))) return rc;
            if ((rc = ops->eri_split(a, cs))) {        // e.g. a launch failure: the thread-per-pair kernel still works
                fprintf(stderr, "[qbx] warp-per-pair Schwarz kernel failed (%s); using one thread per pair\n", qbx_last_error());
                if ((rc = run_eri(pc, pc, t, P.npair, v, cs))) return rc;
            }
        } else if ((rc = run_eri(pc, pc, t, P.npair, v, cs))) return rc;
        const int nab = qbx_nc(P.la) * qbx_nc(P.lb);
        k_schwarz<<<(P.npair + 127) / 128, 128, 0, cs>>>(v, P.npair, nab, P.schwarz);
        QBX_CUDA(cudaGetLastError());
    }
    if ((rc = join(s))) return rc;
    for (void *p : scratch) qbx_pool_free_async(p);          // after the join: later users are ordered behind it
    have_schwarz_ = true;
    return QBX_OK;
}

// Task lists are built in two enqueue-only phases so that a whole store needs two host
// synchronisations instead of several per class:
//   count: per-row counts -> device scan -> the total in d_total[0..2]
//   fill : (the host has read the totals and sized the lists) compaction, statistics into
//          d_stat[0..1] = (valid slots, primitive quartets), chunk order, heavy chunks into d_nheavy.
int Engine::tasks_count(int bc, int kc, double tol, int rank, int nranks, bool grp, TaskScratch &ts, int64_t *d_total,
                        cudaStream_t s)
{
    const DevPairSet &B = pairs_[bc], &K = pairs_[kc];
    ts = TaskScratch();
    if (B.npair == 0 || K.npair == 0) return QBX_OK;
    if (grp) return qbx_group_count(groups_, B, K, bc == kc, tol, rank, nranks, ts, d_total, s);
    QBX_CUDA(qbx_dmalloc(&ts.cnt[0], B.npair * sizeof(int)));
    QBX_CUDA(qbx_dmalloc(&ts.off[0], (B.npair + 1) * sizeof(int64_t)));
    k_count_tasks<<<(unsigned)(((int64_t)B.npair * 32 + 127) / 128), 128, 0, s>>>(B.schwarz, K.schwarz, B.npair, K.npair, bc == kc,
                                                                                 tol, ts.cnt[0]);
    qbx_scan_counts(ts.cnt[0], B.npair, ts.off[0], d_total, s);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

int Engine::tasks_fill(int bc, int kc, double tol, int rank, int nranks, bool grp, bool want_order, TaskScratch &ts,
                       const int64_t *h_total, TaskList &out, double *d_stat, int *d_nheavy, cudaStream_t s)
{
    out = TaskList();
    const DevPairSet &B = pairs_[bc], &K = pairs_[kc];
    int rc = QBX_OK;
    if (B.npair == 0 || K.npair == 0) return QBX_OK;
    if (grp) {
        trace_mark(s, "(before group fill)", bc, kc);
        rc = qbx_group_fill(groups_, B, K, bc == kc, tol, rank, nranks, ts, h_total, out, d_stat, d_nheavy, s);
        trace_mark(s, "group_fill", bc, kc);
    } else {
        const int64_t total = h_total[0];
        const int64_t nfull = total / QBX_TASK_CHUNK, rem = total % QBX_TASK_CHUNK;
        int64_t mine = 0;
        if (nfull > rank) mine = ((nfull - rank + nranks - 1) / nranks) * QBX_TASK_CHUNK;
        if (rem && nfull % nranks == rank) mine += rem;
        out.n = out.nvalid = mine;
        if (mine > 0) {
            QBX_CUDA(qbx_dmalloc(&out.tasks, mine * sizeof(int2)));
            const int64_t threads = (int64_t)B.npair * 32;
            k_fill_tasks<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(B.schwarz, K.schwarz, B.npair, K.npair, bc == kc, tol,
                                                                           ts.off[0], rank, nranks, out.tasks);
            trace_mark(s, "fill_tasks", bc, kc);
            k_task_cost<<<296, 256, 0, s>>>(out.tasks, mine, B.prim_off, K.prim_off, d_stat + 1);
            QBX_CUDA(cudaGetLastError());
            trace_mark(s, "task_cost", bc, kc);
            if (want_order) rc = qbx_chunk_order(out.tasks, nullptr, nullptr, mine, B.prim_off, K.prim_off, &out.order, nullptr, s);
            trace_mark(s, "chunk_order", bc, kc);
        }
    }
    for (int i = 0; i < 3; ++i) { qbx_pool_free_async(ts.cnt[i]); qbx_pool_free_async(ts.off[i]); }
    ts = TaskScratch();
    return rc;
}

// one class on its own (dense-tensor path): both phases with their synchronisations
int Engine::build_tasks(int bc, int kc, double tol, int rank, int nranks, TaskList &out, cudaStream_t s)
{
    out = TaskList();
    if (pairs_[bc].npair == 0 || pairs_[kc].npair == 0) return QBX_OK;
    char *h = (char *)qbx_pinned(64), *d = nullptr;
    if (!h) { qbx_set_error("pinned scratch allocation failed"); return QBX_ERR_NOMEM; }
    QBX_CUDA(qbx_dmalloc(&d, 64));
    QBX_CUDA(cudaMemsetAsync(d, 0, 64, s));
    TaskScratch ts;
    int rc = tasks_count(bc, kc, tol, rank, nranks, false, ts, (int64_t *)d, s);
    if (rc) return rc;
    QBX_CUDA(cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, s));
    QBX_CUDA(cudaStreamSynchronize(s));
    int64_t tot[3];
    memcpy(tot, h, 24);
    if ((rc = tasks_fill(bc, kc, tol, rank, nranks, false, false, ts, tot, out, (double *)(d + 24), nullptr, s))) return rc;
    QBX_CUDA(cudaMemcpyAsync(h, d + 24, 16, cudaMemcpyDeviceToHost, s));
    QBX_CUDA(cudaStreamSynchronize(s));
    out.nprimq = ((double *)h)[1];
    qbx_pool_free_async(d);
    return QBX_OK;
}

static const int64_t kChunkDoubles = 8ll << 20;      // 64 MiB staging block: stays in the 126 MB L2

int Engine::fill_tensor(double *d_tensor, cudaStream_t s, double *stats)
{
    int rc = ensure_schwarz(s);
    if (rc) return rc;
    if (!chunk_) { QBX_CUDA(qbx_dmalloc(&chunk_, kChunkDoubles * sizeof(double))); chunk_doubles_ = kChunkDoubles; }
    for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
        for (int kc = 0; kc <= bc; ++kc) {
            TaskList tl;
            if ((rc = build_tasks(bc, kc, 0.0, 0, 1, tl, s))) return rc;
            if (tl.n == 0) continue;
            const ClassOps *ops = qbx_class_ops(bc, kc);
            const int64_t step = std::max<int64_t>(128, (chunk_doubles_ / ops->ncomp) / 128 * 128);
            for (int64_t o = 0; o < tl.n; o += step) {
                const int64_t n = std::min(step, tl.n - o);
                if ((rc = run_eri(bc, kc, tl.tasks + o, n, chunk_, s))) return rc;
                ScatterArgs a{pairs_[bc].shells, pairs_[kc].shells, tl.tasks + o, n, chunk_, d_shell_bf_, nbf_, d_tensor};
                if ((rc = ops->scatter(a, s))) return rc;
                stats[0] += 2;
            }
            stats[3] += tl.nprimq;
            QBX_CUDA(cudaStreamSynchronize(s));
            qbx_pool_free(tl.tasks);
        }
    return QBX_OK;
}

int Engine::store(double tol, int mode, int rank, int nranks, cudaStream_t s, double *stats)
{
    // QBX_TRACE=1: host-side phase times of this call on stderr (adds stream synchronisations)
    const bool trace = getenv("QBX_TRACE") && atoi(getenv("QBX_TRACE"));
    auto now = [&] { if (trace) cudaStreamSynchronize(s); return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    double t_tasks = 0, t_order = 0;
    auto T0 = now();
    release_store();
    int rc = ensure_schwarz(s);
    if (rc) return rc;
    auto T1 = now();
    // device scratch of the plan: totals[21][3] (int64), stats[21][2] (double), nheavy[21] (int)
    const size_t o_stat = QBX_NCLASS * 3 * sizeof(int64_t), o_nh = o_stat + QBX_NCLASS * 2 * sizeof(double);
    const size_t plan_bytes = o_nh + QBX_NCLASS * sizeof(int);
    char *d_plan = nullptr, *h_plan = (char *)qbx_pinned(plan_bytes);
    if (!h_plan) { qbx_set_error("pinned scratch allocation failed"); return QBX_ERR_NOMEM; }
    QBX_CUDA(qbx_dmalloc(&d_plan, plan_bytes));
    QBX_CUDA(cudaMemsetAsync(d_plan, 0, plan_bytes, s));
    TaskScratch scratch[QBX_NCLASS];
    // The 21 classes are independent and their list kernels are small (a warp per bra row): they go round-robin over
    // the side streams, in both phases, so that the device overlaps them instead of draining ~190 launches one by one.
    // Scratch blocks freed meanwhile are held back until the streams have been joined (QbxPoolDeferScope).
    static const bool trace2 = getenv("QBX_TRACE") && atoi(getenv("QBX_TRACE")) >= 2;     // per-step marks need one stream
    const bool spread = !trace2;
    if (spread && (rc = fork(s))) return rc;
    {
        QbxPoolDeferScope held;
        int c = 0;
        for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
            for (int kc = 0; kc <= bc; ++kc, ++c)
                if ((rc = tasks_count(bc, kc, tol, rank, nranks, mode == 0 && grouped(bc, kc), scratch[c],
                                      (int64_t *)d_plan + 3 * c, spread ? side_[c % kSide] : s))) {
                    if (spread) join(s);
                    return rc;
                }
        if (spread && (rc = join(s))) return rc;
    }
    QBX_CUDA(cudaMemcpyAsync(h_plan, d_plan, o_stat, cudaMemcpyDeviceToHost, s));
    QBX_CUDA(cudaStreamSynchronize(s));                       // host sync 1 of 2: list sizes
    int64_t totals[QBX_NCLASS * 3];
    memcpy(totals, h_plan, o_stat);
    auto T1b = now();
    t_tasks = secs(T1, T1b);
    {
        QbxPoolDeferScope held;
        if (spread && (rc = fork(s))) return rc;
        int c = 0;
        for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
            for (int kc = 0; kc <= bc; ++kc, ++c) {
                if ((rc = tasks_fill(bc, kc, tol, rank, nranks, mode == 0 && grouped(bc, kc), mode == 0, scratch[c], totals + 3 * c,
                                     tasks_[bc][kc], (double *)(d_plan + o_stat) + 2 * c, (int *)(d_plan + o_nh) + c,
                                     spread ? side_[c % kSide] : s))) {
                    if (spread) join(s);
                    return rc;
                }
                const TaskList &tl = tasks_[bc][kc];
                trace_mark(s, "(before vals alloc)", bc, kc);
                if (mode == 0 && tl.n > 0) {
                    const size_t bytes = (size_t)tl.n * qbx_class_ops(bc, kc)->ncomp * sizeof(double);
                    if (qbx_dmalloc(&vals_[bc][kc], bytes) != cudaSuccess) {
                        cudaGetLastError();
                        qbx_set_error("qbx_eri_store: packed ERI store does not fit in device memory; use mode 1 (direct)");
                        release_store();
                        return QBX_ERR_NOMEM;
                    }
                    stored_bytes_ += (int64_t)bytes;
                }
            }
        if (spread && (rc = join(s))) return rc;
        held.release();
    }
    QBX_CUDA(cudaMemcpyAsync(h_plan + o_stat, d_plan + o_stat, plan_bytes - o_stat, cudaMemcpyDeviceToHost, s));
    QBX_CUDA(cudaStreamSynchronize(s));                       // host sync 2 of 2: statistics, heavy-chunk counts
    qbx_pool_free_async(d_plan);
    {
        int c = 0;
        for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
            for (int kc = 0; kc <= bc; ++kc, ++c) {
                TaskList &tl = tasks_[bc][kc];
                const double *st = (const double *)(h_plan + o_stat) + 2 * c;
                if (tl.ngt > 0) { tl.nvalid = (int64_t)st[0]; tl.nheavy = ((const int *)(h_plan + o_nh))[c]; }
                tl.nprimq = st[1];
                const ClassOps *ops = qbx_class_ops(bc, kc);
                n_quartets_ += tl.nvalid;
                n_values_ += tl.nvalid * ops->ncomp;
                n_primq_ += tl.nprimq;
                model_flops_ += tl.nprimq * qbx_model_flops_prim(ops->la, ops->lb, ops->lc, ops->ld) +
                                (double)tl.nvalid * qbx_model_flops_hrr(ops->la, ops->lb, ops->lc, ops->ld);
            }
    }
    t_order = secs(T1b, now());
    auto T2 = now();
    if (!d_Jt_) {
        QBX_CUDA(qbx_dmalloc(&d_Jt_, nint_ * nint_ * sizeof(double)));
        QBX_CUDA(qbx_dmalloc(&d_Kt_, 2 * nint_ * nint_ * sizeof(double)));
        QBX_CUDA(qbx_dmalloc(&d_Dint_, 3 * nint_ * nint_ * sizeof(double)));
    }
    mode_ = mode;
    if (mode == 0) {
        if ((rc = recompute(s, stats))) return rc;
    } else if (!chunk_) {
        QBX_CUDA(qbx_dmalloc(&chunk_, kChunkDoubles * sizeof(double)));
        chunk_doubles_ = kChunkDoubles;
    }
    QBX_CUDA(cudaStreamSynchronize(s));
    if (trace)
        fprintf(stderr, "qbx_eri_store: schwarz %.4f  count %.4f  fill+order+alloc %.4f  eri %.4f s\n", secs(T0, T1), t_tasks,
                t_order, secs(T2, std::chrono::steady_clock::now()));
    return QBX_OK;
}

int Engine::fork(cudaStream_t s)
{
    if (!fork_ev_) {
        QBX_CUDA(cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming));
        for (int i = 0; i < kSide; ++i) {
            QBX_CUDA(cudaStreamCreateWithFlags(&side_[i], cudaStreamNonBlocking));
            QBX_CUDA(cudaEventCreateWithFlags(&side_ev_[i], cudaEventDisableTiming));
        }
    }
    QBX_CUDA(cudaEventRecord(fork_ev_, s));
    for (int i = 0; i < kSide; ++i) QBX_CUDA(cudaStreamWaitEvent(side_[i], fork_ev_, 0));
    return QBX_OK;
}

int Engine::join(cudaStream_t s)
{
    for (int i = 0; i < kSide; ++i) {
        QBX_CUDA(cudaEventRecord(side_ev_[i], side_[i]));
        QBX_CUDA(cudaStreamWaitEvent(s, side_ev_[i], 0));
    }
    return QBX_OK;
}

int Engine::recompute(cudaStream_t s, double *stats, bool timed)
{
    if (mode_ != 0) { qbx_set_error("recompute: stored mode only"); return QBX_ERR_STATE; }
    int rc;
    if (!timed && (rc = fork(s))) return rc;
    int c = 0, k = 0;
    for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
        for (int kc = 0; kc <= bc; ++kc, ++c) {
            if (timed) {
                if (!cls_ev_[c]) QBX_CUDA(cudaEventCreate(&cls_ev_[c]));
                QBX_CUDA(cudaEventRecord(cls_ev_[c], s));
            }
            const TaskList &tl = tasks_[bc][kc];
            if (tl.n == 0) continue;
            cudaStream_t cs = timed ? s : side_[k++ % kSide];
            if (tl.ngt > 0) {
                ClassArgs a;
                if ((rc = eri_args(bc, kc, tl.tasks, tl.n, vals_[bc][kc], cs, a))) return rc;
                rc = qbx_group_eri(kClsLa[bc], groups_, a, tl, cs);
            } else {
                rc = run_eri(bc, kc, tl.tasks, tl.n, vals_[bc][kc], cs, tl.order);
            }
            if (rc) return rc;
            stats[0] += 1;
        }
    if (timed) {
        if (!cls_ev_[QBX_NCLASS]) QBX_CUDA(cudaEventCreate(&cls_ev_[QBX_NCLASS]));
        QBX_CUDA(cudaEventRecord(cls_ev_[QBX_NCLASS], s));
        cls_timed_ = true;
    } else if ((rc = join(s))) return rc;
    stats[3] += n_primq_;
    stats[4] += model_flops_;
    return QBX_OK;
}

int Engine::class_stats(cudaStream_t s, double *stats, double *out)
{
    int rc = recompute(s, stats, true);
    if (rc) return rc;
    QBX_CUDA(cudaEventSynchronize(cls_ev_[QBX_NCLASS]));
    int c = 0;
    for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
        for (int kc = 0; kc <= bc; ++kc, ++c) {
            const ClassOps *ops = qbx_class_ops(bc, kc);
            const TaskList &tl = tasks_[bc][kc];
            float ms = 0;
            QBX_CUDA(cudaEventElapsedTime(&ms, cls_ev_[c], cls_ev_[c + 1]));
            double *o = out + 6 * c;
            o[0] = ops->la * 1000 + ops->lb * 100 + ops->lc * 10 + ops->ld;
            o[1] = ms * 1e-3;
            o[2] = (double)tl.nvalid;
            o[3] = tl.nprimq;
            o[4] = tl.nprimq * qbx_model_flops_prim(ops->la, ops->lb, ops->lc, ops->ld) +
                   (double)tl.nvalid * qbx_model_flops_hrr(ops->la, ops->lb, ops->lc, ops->ld);
            o[5] = (double)tl.nvalid * ops->ncomp;
        }
    return QBX_OK;
}

void Engine::drop_fock_graph()
{
    if (fock_graph_) { cudaGraphExecDestroy(fock_graph_); fock_graph_ = nullptr; }
    fock_key_ = FockKey();
    fock_seen_ = 0;
}

int Engine::fock(int nmat, const double *dDJ, const double *dDK, double *dG, cudaStream_t s, double *stats)
{
    if (mode_ != 0 && mode_ != 1) { qbx_set_error("fock: no ERI representation stored"); return QBX_ERR_STATE; }
    static const bool use_graph = !(getenv("QBX_FOCK_GRAPH") && atoi(getenv("QBX_FOCK_GRAPH")) == 0);
    if (!use_graph || mode_ != 0 || s == nullptr) return fock_enqueue(nmat, dDJ, dDK, dG, s, stats);
    const bool same = fock_key_.nmat == nmat && fock_key_.dj == dDJ && fock_key_.dk == dDK && fock_key_.g == dG && fock_key_.s == s &&
                      fock_key_.gen == store_gen_;
    if (same && fock_graph_) {
        if (cudaGraphLaunch(fock_graph_, s) == cudaSuccess) {
            for (int i = 0; i < 8; ++i) stats[i] += fock_stats_[i];
            stats[6] += 1;                                     // graph launches (qbx_stats)
            return QBX_OK;
        }
        cudaGetLastError();
        drop_fock_graph();                                     // replay refused: fall through to the eager path
    }
    if (!same) {
        drop_fock_graph();
        fock_key_.nmat = nmat; fock_key_.dj = dDJ; fock_key_.dk = dDK; fock_key_.g = dG; fock_key_.s = s; fock_key_.gen = store_gen_;
    }
    if (++fock_seen_ < 2) return fock_enqueue(nmat, dDJ, dDK, dG, s, stats);     // a one-off build is not worth a capture
    // second build with the same buffers: capture it, launch the graph
    if (!d_bad_) QBX_CUDA(qbx_dmalloc(&d_bad_, sizeof(int)));
    if (!fork_ev_) { int rc = fork(s); if (rc) return rc; rc = join(s); if (rc) return rc; }     // side streams exist before the capture
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        fock_seen_ = -(1 << 30);                               // never try again on this key
        return fock_enqueue(nmat, dDJ, dDK, dG, s, stats);
    }
    double captured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int rc = fock_enqueue(nmat, dDJ, dDK, dG, s, captured);
    cudaGraph_t graph = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(s, &graph);
    if (rc != QBX_OK || ee != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        fock_seen_ = -(1 << 30);
        if (rc != QBX_OK) return rc;
        return fock_enqueue(nmat, dDJ, dDK, dG, s, stats);
    }
    const cudaError_t ie = cudaGraphInstantiate(&fock_graph_, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || !fock_graph_) {
        cudaGetLastError();
        fock_graph_ = nullptr;
        fock_seen_ = -(1 << 30);
        return fock_enqueue(nmat, dDJ, dDK, dG, s, stats);
    }
    for (int i = 0; i < 8; ++i) fock_stats_[i] = captured[i];
    QBX_CUDA(cudaGraphLaunch(fock_graph_, s));
    for (int i = 0; i < 8; ++i) stats[i] += fock_stats_[i];
    stats[6] += 1;
    return QBX_OK;
}

int Engine::fock_enqueue(int nmat, const double *dDJ, const double *dDK, double *dG, cudaStream_t s, double *stats)
{
    const int64_t NI2 = nint_ * nint_;
    const unsigned pg = (unsigned)((NI2 + 255) / 256);
    double *DJi = d_Dint_, *DKi = d_Dint_ + NI2;
    if (!d_bad_) QBX_CUDA(qbx_dmalloc(&d_bad_, sizeof(int)));
    QBX_CUDA(cudaMemsetAsync(d_bad_, 0, sizeof(int), s));
    k_permute_in<<<pg, 256, 0, s>>>(nbf_, nint_, d_ext_of_int_, dDJ, DJi, d_bad_);
    for (int m = 0; m < nmat; ++m)
        k_permute_in<<<pg, 256, 0, s>>>(nbf_, nint_, d_ext_of_int_, dDK + m * nbf_ * nbf_, DKi + m * NI2, d_bad_);
    QBX_CUDA(cudaMemsetAsync(d_Jt_, 0, NI2 * sizeof(double), s));
    QBX_CUDA(cudaMemsetAsync(d_Kt_, 0, nmat * NI2 * sizeof(double), s));
    QBX_CUDA(cudaMemsetAsync(dG, 0, nmat * nbf_ * nbf_ * sizeof(double), s));
    stats[0] += 1 + nmat;
    int kside = 0;
    if (mode_ == 0) { int rc = fork(s); if (rc) return rc; }
    for (int bc = 0; bc < QBX_NPAIRCLS; ++bc)
        for (int kc = 0; kc <= bc; ++kc) {
            const TaskList &tl = tasks_[bc][kc];
            if (tl.n == 0) continue;
            const ClassOps *ops = qbx_class_ops(bc, kc);
            DigestArgs a;
            a.bra_info = pairs_[bc].info; a.ket_info = pairs_[kc].info;
            static const int spread = getenv("QBX_DIGEST_SPREAD") ? atoi(getenv("QBX_DIGEST_SPREAD")) : QBX_DIGEST_SPREAD;
            a.spread = spread > 0 ? spread : 1;
            a.nbf = (int)nint_; a.nmat = nmat; a.same_class = (bc == kc);
            a.DJ = DJi; a.DK = DKi; a.Jt = d_Jt_; a.Kt = d_Kt_;
            if (mode_ == 0) {
                a.tasks = tl.tasks; a.ntasks = tl.n; a.vals = vals_[bc][kc];
                cudaStream_t ds = side_[kside++ % kSide];
                // group classes with an s or p bra: one lane per group task (eri_group.cu); QBX_DIGEST_GROUP=0: per quartet
                static const bool by_group = !(getenv("QBX_DIGEST_GROUP") && atoi(getenv("QBX_DIGEST_GROUP")) == 0);
                int rc = (by_group && tl.ngt > 0) ? qbx_group_digest(ops->la, groups_, a, tl, ds) : -1;
                if (rc == -1) rc = ops->digest(a, ds);
                if (rc) return rc;
                stats[0] += 1;
                stats[5] += (double)tl.n * ops->ncomp * sizeof(double);
            } else {
                const int64_t step = std::max<int64_t>(128, (chunk_doubles_ / ops->ncomp) / 128 * 128);
                for (int64_t o = 0; o < tl.n; o += step) {
                    const int64_t n = std::min(step, tl.n - o);
                    int rc = run_eri(bc, kc, tl.tasks + o, n, chunk_, s);
                    if (rc) return rc;
                    a.tasks = tl.tasks + o; a.ntasks = n; a.vals = chunk_;
                    rc = ops->digest(a, s);
                    if (rc) return rc;
                    stats[0] += 2;
                }
                stats[3] += tl.nprimq;
            }
        }
    if (mode_ == 1) stats[4] += model_flops_;
    if (mode_ == 0) { int rc = join(s); if (rc) return rc; }
    k_finish_G<<<pg, 256, 0, s>>>(nint_, nbf_, d_ext_of_int_, nmat, d_Jt_, d_Kt_, dG, d_bad_);
    QBX_CUDA(cudaGetLastError());
    stats[0] += 1;
    return QBX_OK;
}

// ------------------------------------------------------------------ synthetic class batches
int Engine::synthetic(int la, int lb, int lc, int ld, int K, int64_t nq, uint64_t seed, double *secs, double *checksum,
                      double *prim_quartets, int64_t nsample, double *sample_out, double *sample_geom, cudaStream_t s)
{
    if (pair_cls(la, lb) < pair_cls(lc, ld)) { std::swap(la, lc); std::swap(lb, ld); }
    const int bc = pair_cls(la, lb), kc = pair_cls(lc, ld);
    const int64_t np = (int64_t)ceil(sqrt((double)nq));
    // splitmix64: centres uniform in a 10-bohr cube, exponents log-uniform in
    // [0.1, 1e3], coefficients uniform in [-1, 1]
    uint64_t st = seed ^ ((uint64_t)(la * 27 + lb * 9 + lc * 3 + ld) << 32) ^ ((uint64_t)K << 48) ^ 20261017ull;
    auto next = [&]() {
        uint64_t z = (st += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return (double)((z ^ (z >> 31)) >> 11) * (1.0 / 9007199254740992.0);
    };
    std::vector<HostShell> sh;
    auto add = [&](int l) {
        HostShell h;
        h.l = l;
        for (int d = 0; d < 3; ++d) h.cen[d] = 10.0 * next();
        for (int k = 0; k < K; ++k) { h.xpn.push_back(0.1 * pow(1e4, next())); h.coef.push_back(2.0 * next() - 1.0); }
        sh.push_back(h);
    };
    for (int64_t i = 0; i < np; ++i) { add(la); add(lb); }
    for (int64_t i = 0; i < np; ++i) { add(lc); add(ld); }
    // bra pairs come first, ket pairs second; when the pair classes coincide they share one set
    Engine *e = from_shells(sh, 0, true);
    if (!e) return QBX_ERR_CUDA;
    const ClassOps *ops = qbx_class_ops(bc, kc);
    // pair indices: adjacent pairs were appended in order, so within a class the first np are
    // the bra pairs and (if bc == kc) the next np the ket pairs
    const int koff = (bc == kc) ? (int)np : 0;
    std::vector<int2> tasks((size_t)nq);
    for (int64_t q = 0; q < nq; ++q) tasks[q] = make_int2((int)(q / np), koff + (int)(q % np));
    if (prim_quartets) {
        double tot = 0;
        for (int64_t q = 0; q < nq; ++q) tot += (double)e->pairs_[bc].h_nprim[tasks[q].x] * (double)e->pairs_[kc].h_nprim[tasks[q].y];
        *prim_quartets = tot;
    }
    int2 *d_t = nullptr; double *d_v = nullptr, *d_sum = nullptr;
    int rc = QBX_OK;
    do {
        if (qbx_dmalloc(&d_t, nq * sizeof(int2)) != cudaSuccess || qbx_dmalloc(&d_v, (size_t)nq * ops->ncomp * sizeof(double)) != cudaSuccess ||
            qbx_dmalloc(&d_sum, sizeof(double)) != cudaSuccess) { qbx_set_error("qbx_prim_batch: out of device memory"); rc = QBX_ERR_NOMEM; break; }
        cudaMemcpyAsync(d_t, tasks.data(), nq * sizeof(int2), cudaMemcpyHostToDevice, s);
        if ((rc = e->run_eri(bc, kc, d_t, nq, d_v, s))) break;           // warm-up
        cudaEventRecord(e->ev0_, s);
        if ((rc = e->run_eri(bc, kc, d_t, nq, d_v, s))) break;
        cudaEventRecord(e->ev1_, s);
        cudaMemsetAsync(d_sum, 0, sizeof(double), s);
        k_sum<<<296, 256, 0, s>>>(d_v, nq * ops->ncomp, d_sum);
        cudaMemcpyAsync(checksum, d_sum, sizeof(double), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { qbx_set_error(std::string("qbx_prim_batch: ") + cudaGetErrorString(cudaGetLastError())); rc = QBX_ERR_CUDA; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e->ev0_, e->ev1_);
        *secs = ms * 1e-3;
        const int64_t ns = std::min(nsample, nq);
        if (ns > 0 && sample_out && sample_geom) {
            std::vector<double> all((size_t)ops->ncomp);
            for (int64_t q = 0; q < ns; ++q) {
                cudaMemcpy2D(sample_out + q * ops->ncomp, sizeof(double), d_v + q, nq * sizeof(double), sizeof(double),
                             ops->ncomp, cudaMemcpyDeviceToHost);
                const int ids[4] = {2 * (int)(q / np), 2 * (int)(q / np) + 1, 2 * (int)(np + q % np), 2 * (int)(np + q % np) + 1};
                double *g = sample_geom + q * 4 * (3 + 2 * K);
                for (int t = 0; t < 4; ++t) {
                    const HostShell &h = sh[ids[t]];
                    for (int d = 0; d < 3; ++d) g[t * (3 + 2 * K) + d] = h.cen[d];
                    for (int k = 0; k < K; ++k) { g[t * (3 + 2 * K) + 3 + k] = h.xpn[k]; g[t * (3 + 2 * K) + 3 + K + k] = h.coef[k]; }
                }
            }
        }
    } while (0);
    qbx_pool_free(d_t); qbx_pool_free(d_v); qbx_pool_free(d_sum);
    delete e;
    return rc;
}
