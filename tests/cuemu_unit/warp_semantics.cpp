// Unit test of the cuemu shim itself (tools/cuemu): the warp-level primitives must behave as CUDA
// documents them, otherwise a green emulated parity test would prove nothing about the GPU.
// Built and run by tests/test_emu_kernels.py::test_emulator_warp_semantics.  TEST INFRASTRUCTURE.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { ++fails; fprintf(stderr, "FAIL line %d: %s (lane %d)\n", __LINE__, #cond, (int)(threadIdx.x & 31)); } } while (0)

__global__ void k_shuffles(int *out)
{
    const int lane = threadIdx.x & 31;
    const double x = 100.0 + lane;
    CHECK(__shfl_sync(0xffffffffu, x, 5) == 105.0);
    CHECK(__shfl_sync(0xffffffffu, lane, lane ^ 1) == (lane ^ 1));
    CHECK(__shfl_xor_sync(0xffffffffu, x, 16) == 100.0 + (lane ^ 16));
    CHECK(__shfl_up_sync(0xffffffffu, lane, 3) == (lane >= 3 ? lane - 3 : lane));          // out of range: own value
    CHECK(__shfl_down_sync(0xffffffffu, lane, 7) == (lane + 7 <= 31 ? lane + 7 : lane));
    CHECK(__shfl_sync(0xffffffffu, lane, 3, 8) == (lane & ~7) + 3);                        // width = 8: segments of 8 lanes
    const unsigned b = __ballot_sync(0xffffffffu, lane % 3 == 0);
    unsigned expect = 0;
    for (int l = 0; l < 32; l += 3) expect |= 1u << l;
    CHECK(b == expect);
    CHECK(__all_sync(0xffffffffu, lane < 32) == 1);
    CHECK(__all_sync(0xffffffffu, lane < 31) == 0);
    CHECK(__any_sync(0xffffffffu, lane == 17) == 1);
    CHECK(__popc(b) == 11 && __ffs(8) == 4);
    // butterfly sum of doubles: every lane ends with 0 + 1 + ... + 31
    double s = lane;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    CHECK(s == 496.0);
    if (lane == 0) atomicAdd(out, 1);
}

// lanes that have exited do not take part: ballot sees them as 0, the others are still released
__global__ void k_partial(int *out)
{
    const int lane = threadIdx.x & 31;
    if (lane >= 20) return;
    const unsigned b = __ballot_sync(0xffffffffu, 1);
    CHECK(b == 0x000fffffu);
    CHECK(__all_sync(0xffffffffu, lane < 20) == 1);
    __syncwarp();
    if (lane == 0) atomicAdd(out, 1);
}

// shared memory, __syncthreads and blocks: a block-wide reversal through dynamic shared memory
__global__ void k_block(const int *in, int *out, int n)
{
    extern __shared__ int buf[];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    buf[threadIdx.x] = i < n ? in[i] : -1;
    __syncthreads();
    const int j = blockDim.x - 1 - threadIdx.x;
    if (i < n) out[i] = buf[j];
}

int main()
{
    int *cnt = nullptr;
    cudaMalloc(&cnt, sizeof(int));
    cudaMemset(cnt, 0, sizeof(int));
    k_shuffles<<<3, 96>>>(cnt);
    k_partial<<<2, 64>>>(cnt);
    if (*cnt != 3 * 3 + 2 * 2) { ++fails; fprintf(stderr, "FAIL: %d warps reported, expected 13\n", *cnt); }
    const int n = 256, bs = 64;
    std::vector<int> in(n), out(n, 0);
    for (int i = 0; i < n; ++i) in[i] = i;
    int *din = nullptr, *dout = nullptr;
    cudaMalloc(&din, n * sizeof(int)); cudaMalloc(&dout, n * sizeof(int));
    cudaMemcpy(din, in.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    k_block<<<n / bs, bs, bs * sizeof(int)>>>(din, dout, n);
    cudaMemcpy(out.data(), dout, n * sizeof(int), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i)
        if (out[i] != (i / bs) * bs + (bs - 1 - i % bs)) { ++fails; fprintf(stderr, "FAIL: block reversal at %d\n", i); break; }
    printf(fails ? "cuemu warp semantics: %d FAILURES\n" : "cuemu warp semantics: ok\n", fails);
    return fails ? 1 : 0;
}
