// cuemu runtime: coroutine scheduler behind the <<<...>>> launches (see include/cuda_runtime.h).
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <sys/mman.h>
#include <time.h>

#include <mutex>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#endif

#if !defined(__x86_64__)
#error "cuemu's context switch is written for x86-64"
#endif

namespace cuemu {

ThreadCtx *cur = nullptr;

// ---- context switch: callee-saved registers on the coroutine's own stack ----
extern "C" void cuemu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cuemu_switch,.-cuemu_switch
)");

enum State { RUN = 0, WAIT_WARP = 1, WAIT_BLOCK = 2, DONE = 3 };

struct Thread {
    ThreadCtx ctx;
    void *sp;
    int state;
};
struct Warp {
    uint64_t slot[2][32];
    unsigned active[2];
    int parity;
};

static const size_t STACK_BYTES = (size_t)1 << 20;      // per device thread (ERI kernels hold thousands of doubles)
static char *g_stacks = nullptr;
static size_t g_nstacks = 0;
static std::vector<Thread> g_threads;
static std::vector<Warp> g_warps;
static std::vector<char> g_smem;
static void *g_sched_sp = nullptr;
static Thread *g_running = nullptr;
static KernelBody *g_body = nullptr;
static std::mutex g_launch_mutex;
static const bool g_reverse = [] { const char *e = getenv("QBX_EMU_LANE_ORDER"); return e && e[0] == 'r'; }();

int sm_count()
{
    static int n = [] { const char *e = getenv("QBX_EMU_SMS"); int v = e ? atoi(e) : 2; return v > 0 ? v : 2; }();
    return n;
}
double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
void *dyn_smem() { return g_smem.data(); }

static void yield_to_scheduler() { cuemu_switch(&g_running->sp, g_sched_sp); }

void warp_barrier()
{
    g_running->state = WAIT_WARP;
    yield_to_scheduler();
}
void block_barrier()
{
    g_running->state = WAIT_BLOCK;
    yield_to_scheduler();
}
uint64_t *xchg_slot(int parity, int lane) { return &g_warps[cur->warp].slot[parity][lane]; }
int xchg_parity() { return g_warps[cur->warp].parity; }
unsigned active_mask(int parity) { return g_warps[cur->warp].active[parity]; }

static void trampoline()
{
    g_body->run();
    g_running->state = DONE;
    yield_to_scheduler();
    abort();                                    // a finished coroutine is never resumed
}

static void prepare(Thread &t, size_t idx)
{
    char *top = g_stacks + (idx + 1) * STACK_BYTES;
#if defined(__SANITIZE_ADDRESS__)
    // frames of device threads that ended inside the trampoline were never unwound
    __asan_unpoison_memory_region(g_stacks + idx * STACK_BYTES, STACK_BYTES);
#endif
    void **sp = (void **)top;
    *--sp = nullptr;                            // fake return address of trampoline (keeps the ABI alignment)
    *--sp = (void *)&trampoline;
    for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    t.sp = sp;
    t.state = RUN;
}

static void resume(Thread &t)
{
    g_running = &t;
    cur = &t.ctx;
    cuemu_switch(&g_sched_sp, t.sp);
    g_running = nullptr;
    cur = nullptr;
}

void launch_impl(dim3 grid, dim3 block, size_t smem, KernelBody &body)
{
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    const size_t nthr = (size_t)block.x * block.y * block.z;
    const size_t nwarp = (nthr + 31) / 32;
    if (nthr == 0 || (size_t)grid.x * grid.y * grid.z == 0) return;
    if (nthr > g_nstacks) {
        if (g_stacks) munmap(g_stacks, g_nstacks * STACK_BYTES);
        g_stacks = (char *)mmap(nullptr, nthr * STACK_BYTES, PROT_READ | PROT_WRITE,
                                MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == (char *)MAP_FAILED) { fprintf(stderr, "cuemu: cannot map coroutine stacks\n"); abort(); }
        g_nstacks = nthr;
    }
    g_threads.resize(nthr);
    g_warps.resize(nwarp);
    if (g_smem.size() < smem + 64) g_smem.resize(smem + 64);
    g_body = &body;

    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        for (size_t i = 0; i < nthr; ++i) {
            Thread &t = g_threads[i];
            t.ctx.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / ((size_t)block.x * block.y))};
            t.ctx.bid = uint3{bx, by, bz};
            t.ctx.bdim = block;
            t.ctx.gdim = grid;
            t.ctx.lane = (int)(i & 31);
            t.ctx.warp = (int)(i >> 5);
            prepare(t, i);
        }
        for (auto &w : g_warps) { w.parity = 0; w.active[0] = w.active[1] = 0; }
        for (;;) {
            bool progress = false;
            size_t live_block = 0, wait_block = 0;
            for (size_t wn = 0; wn < nwarp; ++wn) {
                const size_t w = g_reverse ? nwarp - 1 - wn : wn;          // warps too: missing __syncthreads
                const size_t lo = w * 32, hi = lo + 32 < nthr ? lo + 32 : nthr;
                for (;;) {
                    // lane order inside a warp: ascending, or descending with QBX_EMU_LANE_ORDER=reverse.
                    // Code that is correct under both orders has no missing __syncwarp between a lane's
                    // write and another lane's read of the same shared location (either direction).
                    for (size_t n = 0; n < hi - lo; ++n) {
                        const size_t i = g_reverse ? hi - 1 - n : lo + n;
                        if (g_threads[i].state == RUN) { resume(g_threads[i]); progress = true; }
                    }
                    unsigned live = 0, waiting = 0, mask = 0;
                    for (size_t i = lo; i < hi; ++i) {
                        const int s = g_threads[i].state;
                        if (s != DONE) ++live;
                        if (s == WAIT_WARP) { ++waiting; mask |= 1u << (i - lo); }
                    }
                    if (live > 0 && waiting == live) {          // release the warp
                        Warp &W = g_warps[w];
                        W.active[W.parity] = mask;
                        W.parity ^= 1;
                        for (size_t i = lo; i < hi; ++i)
                            if (g_threads[i].state == WAIT_WARP) g_threads[i].state = RUN;
                        continue;
                    }
                    break;
                }
                for (size_t i = lo; i < hi; ++i) {
                    const int s = g_threads[i].state;
                    if (s != DONE) ++live_block;
                    if (s == WAIT_BLOCK) ++wait_block;
                }
            }
            if (live_block == 0) break;
            if (wait_block == live_block) {
                for (auto &t : g_threads)
                    if (t.state == WAIT_BLOCK) t.state = RUN;
                continue;
            }
            if (!progress) {
                fprintf(stderr, "cuemu: DEADLOCK in block (%u,%u,%u): %zu live threads, %zu at __syncthreads, the rest "
                                "at a warp-level sync that not all live lanes of their warp reach\n",
                        bx, by, bz, live_block, wait_block);
                abort();
            }
        }
    }
    g_body = nullptr;
}

}   // namespace cuemu
