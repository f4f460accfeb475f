// K4: GroupNorm32 (+SiLU) on channels-last bf16 activations, fp32 math (guided_diffusion/nn.py:17-19).
//
// Two HBM-bound passes: `stats` reads x once (per-thread fp32 partial sums -> shared-memory atomics ->
// one fp64 atomicAdd per (block, group)), `apply` reads x once and writes y once with the affine
// transform and SiLU fused.  A group at full latent resolution is 2 channels x 1,003,520 voxels, far
// larger than shared memory, hence the split.
#include "common.cuh"

namespace fcwdm {

constexpr int kNormThreads = 256;

// Replicas of the (sum, sumsq) accumulators: block b adds into replica b % kStatReplicas, so at most
// gridDim.x / kStatReplicas fp64 atomics hit one L2 address (same-address atomics serialise in the L2 slice).
constexpr int kStatReplicas = 16;

__global__ void __launch_bounds__(kNormThreads) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld,
                                                                double* __restrict__ stats, int64_t S, int C, int G) {
    pdl_prologue();
    extern __shared__ float sm[];  // [16][kNormThreads] per-thread partials, then [2][C] per-channel sums
    float* part = sm;
    float* chs = sm + 16 * kNormThreads;
    const int C8 = C >> 3;
    const int n = blockIdx.y;
    const int nthr = blockDim.x;           // the largest multiple of C/8 that fits kNormThreads
    const int chunk = threadIdx.x % C8;
    const int lane = threadIdx.x / C8;
    const int vpb = nthr / C8;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const __nv_bfloat16* base = x + (int64_t)n * S * ld + chunk * 8;
    const int64_t stride = (int64_t)gridDim.x * vpb;
    int64_t v = (int64_t)blockIdx.x * vpb + lane;
    // 4 independent 128-bit loads in flight per thread (memory-level parallelism), then the arithmetic
    for (; v + 3 * stride < S; v += 4 * stride) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = ld_stream_u4(base + (v + k * stride) * ld);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += f[j];
                q[j] = fmaf(f[j], f[j], q[j]);
            }
        }
    }
    for (; v < S; v += stride) {
        float f[8];
        unpack8(ld_stream_u4(base + v * ld), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] += f[j];
            q[j] = fmaf(f[j], f[j], q[j]);
        }
    }
    // deterministic in-block reduction: fixed summation order over the vpb voxel lanes of each channel
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        part[j * kNormThreads + threadIdx.x] = s[j];
        part[(8 + j) * kNormThreads + threadIdx.x] = q[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += nthr) {
        const int comp = i / C, c = i % C;
        const float* p = part + (comp * 8 + (c & 7)) * kNormThreads + (c >> 3);
        float a = 0.f;
        for (int l = 0; l < vpb; ++l) a += p[l * C8];
        chs[i] = a;
    }
    __syncthreads();
    const int cpg = C / G;
    double* dst = stats + (((int64_t)n * kStatReplicas + (blockIdx.x % kStatReplicas)) * G) * 2;
    for (int g = threadIdx.x; g < G; g += nthr) {
        double a = 0.0, b = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            a += (double)chs[c];
            b += (double)chs[C + c];
        }
        atomicAdd(&dst[g * 2 + 0], a);
        atomicAdd(&dst[g * 2 + 1], b);
    }
}

// sigmoid through the hardware tanh (one MUFU op): silu(x) = x * (0.5 * tanh(x / 2) + 0.5); relative error ~2^-11,
// far below the bf16 output rounding
__device__ __forceinline__ float silu_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return x * fmaf(0.5f, t, 0.5f);
}

template <bool kSilu>
__global__ void __launch_bounds__(kNormThreads) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld,
                                                                __nv_bfloat16* __restrict__ y, int64_t y_ld,
                                                                const double* __restrict__ stats,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int64_t S, int C, int G,
                                                                float eps) {
    pdl_prologue();
    extern __shared__ float sm[];  // [2][C]: scale, shift
    const int n = blockIdx.y;
    const int cpg = C / G;
    const double cnt = (double)S * (double)cpg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        double sum = 0.0, sq = 0.0;
        for (int r = 0; r < kStatReplicas; ++r) {
            sum += stats[(((int64_t)n * kStatReplicas + r) * G + g) * 2 + 0];
            sq += stats[(((int64_t)n * kStatReplicas + r) * G + g) * 2 + 1];
        }
        const double mean = sum / cnt;
        double var = sq / cnt - mean * mean;  // biased variance, as nn.GroupNorm
        var = var < 0.0 ? 0.0 : var;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = rstd * gamma[c];
        sm[c] = sc;
        sm[C + c] = beta[c] - (float)mean * sc;
    }
    __syncthreads();
    const int C8 = C >> 3;
    const int chunk = threadIdx.x % C8;
    const int lane = threadIdx.x / C8;
    const int vpb = blockDim.x / C8;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = sm[chunk * 8 + j];
        sh[j] = sm[C + chunk * 8 + j];
    }
    const __nv_bfloat16* xb = x + (int64_t)n * S * x_ld + chunk * 8;
    __nv_bfloat16* yb = y + (int64_t)n * S * y_ld + chunk * 8;
    const int64_t stride = (int64_t)gridDim.x * vpb;
    int64_t v = (int64_t)blockIdx.x * vpb + lane;
    for (; v + 3 * stride < S; v += 4 * stride) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = ld_stream_u4(xb + (v + k * stride) * x_ld);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = fmaf(f[j], sc[j], sh[j]);
                f[j] = kSilu ? silu_fast(t) : t;
            }
            *reinterpret_cast<uint4*>(yb + (v + k * stride) * y_ld) = pack8(f);
        }
    }
    for (; v < S; v += stride) {
        float f[8];
        unpack8(ld_stream_u4(xb + v * x_ld), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = fmaf(f[j], sc[j], sh[j]);
            f[j] = kSilu ? silu_fast(t) : t;
        }
        *reinterpret_cast<uint4*>(yb + v * y_ld) = pack8(f);
    }
}

}  // namespace fcwdm

using namespace fcwdm;

static int gn_check(const char* fn, int64_t N, int64_t S, int64_t C, int64_t G) {
    FCWDM_REQUIRE(N >= 0 && S >= 0 && C > 0 && G > 0 && N <= 65535, FCWDM_ERR_INVALID, "%s: bad dimension", fn);
    FCWDM_REQUIRE(C % G == 0, FCWDM_ERR_INVALID, "%s: C (%lld) not divisible by G (%lld)", fn, (long long)C, (long long)G);
    FCWDM_REQUIRE(C % 8 == 0 && C <= 2048, FCWDM_ERR_UNSUPPORTED, "%s: C must be a multiple of 8, at most 2048 (got %lld)",
                  fn, (long long)C);
    return FCWDM_OK;
}

extern "C" int fcwdm_groupnorm_stats(const void* x, int64_t ld, double* stats, int64_t N, int64_t S, int64_t C,
                                     int64_t G, void* stream) {
    FCWDM_REQUIRE(x && stats, FCWDM_ERR_INVALID, "fcwdm_groupnorm_stats: null pointer");
    int rc = gn_check("fcwdm_groupnorm_stats", N, S, C, G);
    if (rc) return rc;
    FCWDM_REQUIRE(ld >= C && ld % 8 == 0, FCWDM_ERR_INVALID, "fcwdm_groupnorm_stats: bad ld");
    if (N == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * N * G * kStatReplicas, st);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_groupnorm_stats: memset failed (%s)", cudaGetErrorString(e));
    if (S == 0) return FCWDM_OK;
    const int64_t vpb = kNormThreads / (C / 8);
    const unsigned nthr = (unsigned)(vpb * (C / 8));  // 256, or e.g. 240 for C = 192 / 384 (the plain U-Net's concatenated skips)
    int64_t blocks = (S + vpb * 4 - 1) / (vpb * 4);  // >= 4 voxels per thread
    const int64_t cap = (int64_t)num_sms() * 8 / (N > 8 ? 8 : N);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, (unsigned)N);
    launch_k(gn_stats_kernel, dim3(grid), dim3(nthr), (16 * kNormThreads + 2 * C) * sizeof(float), st, (const __nv_bfloat16*)x, ld, stats, S, (int)C,
                                                                        (int)G);
    FCWDM_CHECK_LAUNCH("fcwdm_groupnorm_stats");
    return FCWDM_OK;
}

extern "C" int fcwdm_groupnorm_apply(const void* x, int64_t x_ld, void* y, int64_t y_ld, const double* stats,
                                     const float* gamma, const float* beta, int64_t N, int64_t S, int64_t C, int64_t G,
                                     float eps, int silu, void* stream) {
    FCWDM_REQUIRE(x && y && stats && gamma && beta, FCWDM_ERR_INVALID, "fcwdm_groupnorm_apply: null pointer");
    int rc = gn_check("fcwdm_groupnorm_apply", N, S, C, G);
    if (rc) return rc;
    FCWDM_REQUIRE(x_ld >= C && y_ld >= C && x_ld % 8 == 0 && y_ld % 8 == 0, FCWDM_ERR_INVALID,
                  "fcwdm_groupnorm_apply: bad ld");
    if (N * S == 0) return FCWDM_OK;
    const int64_t vpb = kNormThreads / (C / 8);
    const unsigned nthr = (unsigned)(vpb * (C / 8));
    int64_t blocks = (S + vpb * 4 - 1) / (vpb * 4);
    const int64_t cap = (int64_t)num_sms() * 8 / (N > 8 ? 8 : N);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, (unsigned)N);
    cudaStream_t st = (cudaStream_t)stream;
    if (silu)
        launch_k(gn_apply_kernel<true>, dim3(grid), dim3(nthr), 2 * C * sizeof(float), st, 
            (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, stats, gamma, beta, S, (int)C, (int)G, eps);
    else
        launch_k(gn_apply_kernel<false>, dim3(grid), dim3(nthr), 2 * C * sizeof(float), st, 
            (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, stats, gamma, beta, S, (int)C, (int)G, eps);
    FCWDM_CHECK_LAUNCH("fcwdm_groupnorm_apply");
    return FCWDM_OK;
}
