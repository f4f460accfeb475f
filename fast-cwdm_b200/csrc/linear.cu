// Timestep path: sinusoidal embedding (guided_diffusion/nn.py:103-121) and the small dense layers
// time_embed / emb_layers (guided_diffusion/wunet.py:472-475, 203-206).  Tiny (N x 256) problems: one warp
// per output element, fp32 throughout, launch-latency bound.
#include "common.cuh"

namespace fcwdm {

template <typename T>      // int64 timesteps, or float ones (rescale_timesteps: t * 1000 / T, respace.py:128-132)
__global__ void timestep_embedding_kernel(const T* __restrict__ t, float* __restrict__ out, int64_t N, int dim,
                                          float max_period) {
    pdl_prologue();
    const int half = dim / 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * dim) return;
    const int64_t n = idx / dim;
    const int j = (int)(idx % dim);
    float v = 0.f;  // odd dim: trailing zero column (nn.py:119-120)
    if (j < 2 * half) {
        const int k = j < half ? j : j - half;
        // exp evaluated in fp64 and rounded once: the correctly rounded fp32 value of the reference's fp32 exp argument
        const float freq = (float)exp((double)(-logf(max_period) * (float)k / (float)half));
        const float arg = (float)t[n] * freq;
        v = j < half ? cosf(arg) : sinf(arg);   // cat([cos, sin]) (nn.py:118)
    }
    out[idx] = v;
}

__device__ __forceinline__ float act(float v, int kind) { return kind == 1 ? v / (1.0f + expf(-v)) : v; }

__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                     const float* __restrict__ b, float* __restrict__ y, int64_t N,
                                                     int64_t K, int64_t M, int act_in, int act_out) {
    pdl_prologue();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= N * M) return;
    const int64_t n = warp / M, m = warp % M;
    float acc = 0.f;
    for (int64_t k = lane; k < K; k += 32) acc = fmaf(act(x[n * K + k], act_in), W[m * K + k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[n * M + m] = act(acc + (b ? b[m] : 0.f), act_out);
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_timestep_embedding(const int64_t* t, float* out, int64_t N, int64_t dim, float max_period,
                                        void* stream) {
    FCWDM_REQUIRE(t && out, FCWDM_ERR_INVALID, "fcwdm_timestep_embedding: null pointer");
    FCWDM_REQUIRE(N >= 0 && dim > 0 && max_period > 0.f, FCWDM_ERR_INVALID, "fcwdm_timestep_embedding: bad argument");
    if (N == 0) return FCWDM_OK;
    const int64_t total = N * dim;
    launch_k(timestep_embedding_kernel<int64_t>, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, t, out, N,
             (int)dim, max_period);
    FCWDM_CHECK_LAUNCH("fcwdm_timestep_embedding");
    return FCWDM_OK;
}

extern "C" int fcwdm_timestep_embedding_f32(const float* t, float* out, int64_t N, int64_t dim, float max_period,
                                            void* stream) {
    FCWDM_REQUIRE(t && out, FCWDM_ERR_INVALID, "fcwdm_timestep_embedding_f32: null pointer");
    FCWDM_REQUIRE(N >= 0 && dim > 0 && max_period > 0.f, FCWDM_ERR_INVALID, "fcwdm_timestep_embedding_f32: bad argument");
    if (N == 0) return FCWDM_OK;
    const int64_t total = N * dim;
    launch_k(timestep_embedding_kernel<float>, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, t, out, N,
             (int)dim, max_period);
    FCWDM_CHECK_LAUNCH("fcwdm_timestep_embedding_f32");
    return FCWDM_OK;
}

extern "C" int fcwdm_linear(const float* x, const float* W, const float* b, float* y, int64_t N, int64_t K, int64_t M,
                            int act_in, int act_out, void* stream) {
    FCWDM_REQUIRE(x && W && y, FCWDM_ERR_INVALID, "fcwdm_linear: null pointer");
    FCWDM_REQUIRE(N >= 0 && K > 0 && M >= 0, FCWDM_ERR_INVALID, "fcwdm_linear: bad dimension");
    FCWDM_REQUIRE((act_in == 0 || act_in == 1) && (act_out == 0 || act_out == 1), FCWDM_ERR_INVALID,
                  "fcwdm_linear: activation must be 0 (identity) or 1 (SiLU)");
    if (N * M == 0) return FCWDM_OK;
    const int64_t warps = N * M;
    launch_k(linear_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, W, b, y, N, K, M, act_in,
                                                                                         act_out);
    FCWDM_CHECK_LAUNCH("fcwdm_linear");
    return FCWDM_OK;
}
