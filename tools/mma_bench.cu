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

// the general conv kernel's issue structure: rounds of `per_round` MMAs, each round = wait on a (ready) mbarrier, fence,
// elect, MMAs, commit to another mbarrier, warp sync -- which part costs the ~30 cycles per MMA seen in the real kernel?
__global__ void __launch_bounds__(128, 1) bench_rounds(int N, int rounds, int per_round, int mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar, ready, sink;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        mbar_init(smem_u32(&ready), 1);
        mbar_init(smem_u32(&sink), 1);
        fence_barrier_init();
        mbar_arrive(smem_u32(&ready));            // phase 0 of `ready` is complete: waits on parity 0 return at once
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a0 = make_sw128_desc(base, 1280);
        const uint64_t b0 = make_sw128_desc(base + 96 * 1024, 1024);
        const long long t0 = clock64();
        if (mode & 16) {
            // ONE elected region around all rounds: the barrier wait and the commit are executed by the elected thread
            if (elect_one()) {
                for (int r = 0; r < rounds; ++r) {
                    mbar_wait(smem_u32(&ready), 0);
                    tc_fence_after();
                    for (int i = 0; i < per_round; ++i)
                        umma_bf16(tmem + (uint32_t)((i & 1) * N), a0 + 2 * (i & 3) + 80 * (i % 3), b0 + 2 * (i & 3), idesc, 1u);
                    umma_commit(smem_u32(&sink));
                }
            }
            __syncwarp();
        } else
        for (int r = 0; r < rounds; ++r) {
            if (mode & 1) mbar_wait(smem_u32(&ready), 0);
            if (mode & 2) tc_fence_after();
            if (elect_one()) {
                for (int i = 0; i < per_round; ++i)
                    umma_bf16(tmem + (uint32_t)((i & 1) * N), a0 + 2 * (i & 3) + 80 * (i % 3), b0 + 2 * (i & 3), idesc, 1u);
                if (mode & 4) umma_commit(smem_u32(&sink));
            }
            if (mode & 8) __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if (threadIdx.x == 32) out[0] = clock64() - t0;
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
    cudaFuncSetAttribute(bench_rounds, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"MMAs only", "+ mbarrier wait (ready)", "+ wait + fence", "+ commit per round", "+ warp sync", "all (wait, fence, commit, sync)", "one elect region around all rounds"};
    const int modes[] = {0, 1, 3, 4, 8, 15, 16};
    for (int N : {64, 128}) {
        for (int per : {12, 24}) {
            for (int k = 0; k < 7; ++k) {
                const int rounds = 4096 / per;
                bench_rounds<<<1, 128, 200 * 1024>>>(N, rounds, per, modes[k], d);
                long long h[2];
                if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
                printf("N%-3d %2d MMAs/round  %-34s %6.1f cyc/MMA\n", N, per, names[k], (double)h[0] / (rounds * per));
            }
        }
    }
    return 0;
}
