// cuemu -- a minimal functional emulator of the CUDA execution model on the host CPU.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It lets the *unmodified* kernel sources of
// quiqbox.jl_b200/csrc be compiled with g++ into tests/_emu/libqbx_emu.so, so that the CPU test
// tier (`pytest -m "not gpu"`, no GPU in the build container) can run the kernels' logic --
// recurrences, task lists, work queues, warp shuffles, shared-memory staging, atomics -- against
// the oracle on small molecules.  Nothing under quiqbox.jl_b200/ loads or links it; the product
// library is libqbx.so built by nvcc for sm_100a and fails loudly without a GPU.
//
// This header shadows <cuda_runtime.h> (tools/cuemu/include is first on the include path).
// Execution model:
//   * a kernel launch runs synchronously: blocks one after the other, the threads of a block as
//     coroutines on one host thread; a thread runs until it reaches a warp-level or block-level
//     synchronisation point (__shfl*_sync, __ballot_sync, __all_sync, __syncwarp, __syncthreads),
//     then the next lane runs; a warp (block) is released when all its live lanes have arrived;
//   * a kernel that can never be released (divergent barrier) aborts with a message instead of
//     hanging -- a class of bug the GPU would punish with a hang and a strike;
//   * streams and events are no-ops (everything is synchronous), device memory is host memory,
//     atomics are plain read-modify-writes (one host thread executes all device code).
// The emulated device reports QBX_EMU_SMS multiprocessors (default 2) so persistent grids stay small.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <type_traits>

#define CUEMU 1

// ------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__
#define CUEMU_STATIC_SHARED static

// ------------------------------------------------------------------ vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) int2 { int x, y; };          // CUDA's alignments: a misaligned vector access is
struct alignas(16) int4 { int x, y, z, w; };   // an error on the GPU; UBSan reports it here
struct alignas(8) float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

// ------------------------------------------------------------------ runtime (cuemu_rt.cpp)
namespace cuemu {
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    int lane, warp;
};
extern ThreadCtx *cur;                 // the device thread that is running
void *dyn_smem();
void warp_barrier();                   // returns when all live lanes of the warp have arrived
void block_barrier();
// exchange slots of the current warp: slot(parity, lane); parity() flips at every release
uint64_t *xchg_slot(int parity, int lane);
int xchg_parity();
unsigned active_mask(int parity);      // lanes that took part in the barrier of that parity
struct KernelBody {
    virtual void run() = 0;
    virtual ~KernelBody() {}
};
void launch_impl(dim3 grid, dim3 block, size_t smem, KernelBody &body);
template <class F>
struct Body : KernelBody {
    F f;
    explicit Body(F &&f_) : f(static_cast<F &&>(f_)) {}
    void run() override { f(); }
};
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F &&f)
{
    Body<F> b(static_cast<F &&>(f));
    launch_impl(grid, block, smem, b);
}
}   // namespace cuemu

#define threadIdx (cuemu::cur->tid)
#define blockIdx (cuemu::cur->bid)
#define blockDim (cuemu::cur->bdim)
#define gridDim (cuemu::cur->gdim)
#define warpSize 32

// ------------------------------------------------------------------ device intrinsics
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline void __syncthreads() { cuemu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::warp_barrier(); }
static inline void __threadfence() {}
static inline int __double2int_rn(double x) { return (int)nearbyint(x); }
static inline int __double2int_rd(double x) { return (int)floor(x); }
static inline int __double2int_rz(double x) { return (int)x; }
static inline double __int2double_rn(int x) { return (double)x; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
// CUDA's global min/max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }

template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

namespace cuemu {
template <class T> inline T shfl_from(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    const int p = xchg_parity();
    uint64_t w = 0;
    memcpy(&w, &v, sizeof(T));
    *xchg_slot(p, cur->lane) = w;
    warp_barrier();
    const bool live = (active_mask(p) >> (src_lane & 31)) & 1u;
    w = *xchg_slot(p, live ? (src_lane & 31) : cur->lane);
    T r;
    memcpy(&r, &w, sizeof(T));
    return r;
}
}   // namespace cuemu
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const int base = cuemu::cur->lane & ~(width - 1);
    return cuemu::shfl_from(v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32)
{
    return cuemu::shfl_from(v, cuemu::cur->lane ^ m);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32)
{
    const int s = cuemu::cur->lane - (int)d;
    return cuemu::shfl_from(v, s < 0 ? cuemu::cur->lane : s);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32)
{
    const int s = cuemu::cur->lane + (int)d;
    return cuemu::shfl_from(v, s > 31 ? cuemu::cur->lane : s);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    const int p = cuemu::xchg_parity();
    *cuemu::xchg_slot(p, cuemu::cur->lane) = pred ? 1u : 0u;
    cuemu::warp_barrier();
    const unsigned act = cuemu::active_mask(p);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((act >> l) & 1u) && *cuemu::xchg_slot(p, l)) r |= 1u << l;
    return r;
}
static inline int __all_sync(unsigned m, int pred)
{
    const int p = cuemu::xchg_parity();
    const unsigned b = __ballot_sync(m, pred);
    return b == cuemu::active_mask(p);
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __reduce_min_sync(unsigned, int v)
{
    const int p = cuemu::xchg_parity();
    *cuemu::xchg_slot(p, cuemu::cur->lane) = (uint64_t)(int64_t)v;
    cuemu::warp_barrier();
    const unsigned act = cuemu::active_mask(p);
    int r = 0x7fffffff;
    for (int l = 0; l < 32; ++l)
        if ((act >> l) & 1u) { const int x = (int)(int64_t)*cuemu::xchg_slot(p, l); if (x < r) r = x; }
    return r;
}
// lanes (among the live ones) that hold the same value as the caller
template <class T> static inline unsigned __match_any_sync(unsigned, T v)
{
    static_assert(sizeof(T) <= 8, "match of a type wider than 64 bits");
    const int p = cuemu::xchg_parity();
    uint64_t w = 0;
    memcpy(&w, &v, sizeof(T));
    *cuemu::xchg_slot(p, cuemu::cur->lane) = w;
    cuemu::warp_barrier();
    const unsigned act = cuemu::active_mask(p);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((act >> l) & 1u) && *cuemu::xchg_slot(p, l) == w) r |= 1u << l;
    return r;
}

// ------------------------------------------------------------------ host API
typedef int cudaError_t;
typedef struct cuemuStream *cudaStream_t;
typedef struct cuemuEvent *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount, major, minor, clockRate;
    size_t totalGlobalMem;
};
namespace cuemu { int sm_count(); double now_ms(); }

static inline const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "cuemu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = cuemu::sm_count(); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "cuemu (host emulation)");
    p->multiProcessorCount = cuemu::sm_count();
    p->major = 10; p->minor = 0; p->clockRate = 1000000;
    p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = nullptr;
    if (posix_memalign(p, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind)
{
    for (size_t r = 0; r < h; ++r) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
struct cuemuEvent { double t; };
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuemuEvent{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = cuemu::now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
// stream capture / graphs: not emulated -- capture always fails and the library runs the eager path
typedef struct cuemuGraph *cudaGraph_t;
typedef struct cuemuGraphExec *cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { *g = nullptr; return cudaErrorInvalidValue; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t, unsigned long long = 0) { *e = nullptr; return cudaErrorInvalidValue; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
struct cudaFuncAttributes { size_t localSizeBytes, sharedSizeBytes; int numRegs, maxThreadsPerBlock; };
template <class K> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, K) { memset(a, 0, sizeof(*a)); if (getenv("QBX_EMU_FAKE_SPILL")) a->localSizeBytes = 1024; return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }
