// Development microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16, SS mode, K-major SWIZZLE_128B operands)
// for a few (M, N) shapes, A stride patterns and accumulator patterns.  Operands are whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fast-cwdm_b200/csrc -o tools/_bin/mma_bench tools/mma_bench.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace fcwdm;

__global__ void __launch_bounds__(128, 1) bench(int M, int N, int iters, int n_acc, uint32_t sbo_a, uint32_t a_shift, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t a0 = make_sw128_desc(base + a_shift, sbo_a);
        const uint64_t b0 = make_sw128_desc(base + 96 * 1024, 1024);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t acc = tmem + (uint32_t)((i % n_acc) * N);
                umma_bf16(acc, a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc, 1u);
            }
            umma_commit(smem_u32(&bar));
            t1 = clock64();
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if (elect_one()) { out[0] = t1 - t0; out[1] = clock64() - t0; }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    struct Case { int M, N, n_acc; uint32_t sbo, shift; const char* what; };
    const Case cases[] = {
        {128, 64, 1, 1024, 0, "M128 N64  1 acc aligned"},   {128, 64, 4, 1024, 0, "M128 N64  4 acc aligned"},
        {128, 128, 1, 1024, 0, "M128 N128 1 acc aligned"},  {128, 128, 2, 1024, 0, "M128 N128 2 acc aligned"},
        {128, 256, 1, 1024, 0, "M128 N256 1 acc aligned"},  {128, 256, 2, 1024, 0, "M128 N256 2 acc aligned"},
        {128, 64, 4, 1280, 0, "M128 N64  4 acc pitch 10"},  {128, 64, 4, 1280, 1408, "M128 N64  4 acc pitch 10 shifted 11 rows"},
        {128, 128, 2, 1280, 1408, "M128 N128 2 acc pitch 10 shifted"}, {128, 192, 2, 1280, 1408, "M128 N192 2 acc pitch 10 shifted"},
        {64, 64, 4, 1024, 0, "M64  N64  4 acc aligned"},    {64, 128, 2, 1024, 0, "M64  N128 2 acc aligned"},
        {128, 16, 4, 1024, 0, "M128 N16  4 acc aligned"},   {128, 32, 4, 1024, 0, "M128 N32  4 acc aligned"},
    };
    for (const Case& c : cases) {
        bench<<<1, 128, 200 * 1024>>>(c.M, c.N, iters, c.n_acc, c.sbo, c.shift, d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.what, cudaGetErrorString(e)); return 1; }
        const double flop = 2.0 * c.M * c.N * 16;
        printf("%-44s issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  -> %6.0f flop/clk/SM\n", c.what, (double)h[0] / iters,
               (double)h[1] / iters, flop / ((double)h[1] / iters));
    }
    return 0;
}
