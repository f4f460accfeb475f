// K3: fused reverse-diffusion step, q_sample, final image-space step, and the planar<->channels-last
// converters at the denoiser boundary.  All are single-pass, HBM-bound elementwise kernels: the 2x2x2 Haar
// brick that IDWT -> clamp -> DWT touches is exactly the 8 band values of ONE latent voxel, so the whole
// p_sample update is local (guided_diffusion/gaussian_diffusion.py:335-355, 244-267, 565-573).
#include "common.cuh"

namespace fcwdm {

struct StepCoef {
    float c1, c2, sigma, recip, recipm1;
};

__device__ __forceinline__ void step_voxel(float* mo, const float* xt, const float* nz, const StepCoef& k,
                                           int clip, int predict_xstart, float* xprev, float* pred) {
    float x0[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) x0[b] = predict_xstart ? mo[b] : (k.recip * xt[b] - k.recipm1 * mo[b]);
    if (clip) {
        float bands[8], img[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) bands[b] = x0[b];
        bands[0] *= 3.0f;                       // x[:, 0] * 3.   (:340)
        haar_synthesis(bands, img);
#pragma unroll
        for (int q = 0; q < 8; ++q) img[q] = fminf(fmaxf(img[q], 0.0f), 1.0f);   // clamp(0., 1.) (:349)
        haar_analysis(img, x0);
        x0[0] = x0[0] / 3.0f;                   // LLL / 3.       (:352)
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        pred[b] = x0[b];
        float mean = k.c1 * x0[b] + k.c2 * xt[b];     // q_posterior_mean_variance (:253-256)
        xprev[b] = mean + k.sigma * nz[b];            // p_sample (:573); sigma already carries (t != 0)
    }
}

// One thread = 4 consecutive latent voxels along w (float4 per band plane).  Requires S % 4 == 0.
template <bool kModelOutCl>
__global__ void __launch_bounds__(256) p_sample_step_kernel(const void* __restrict__ model_out, int64_t mo_ld,
                                                            const float* x_t,      // may alias x_prev (in-place update)
                                                            const float* __restrict__ noise, float* x_prev,
                                                            float* __restrict__ pred_xstart,
                                                            __nv_bfloat16* __restrict__ x_prev_cl, int64_t xp_ld,
                                                            const float* __restrict__ coef,
                                                            const int64_t* __restrict__ t, int64_t T, int64_t N,
                                                            int64_t S, int clip, int predict_xstart) {
    pdl_prologue();
    const int64_t S4 = S >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S4) return;
    const int64_t n = idx / S4;
    const int64_t s = (idx % S4) * 4;
    int64_t ti = t[n];
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);  // range is validated on the host (IndexError in the reference, :1257)
    StepCoef k;
    k.c1 = coef[ti * 5 + 0]; k.c2 = coef[ti * 5 + 1]; k.sigma = coef[ti * 5 + 2];
    k.recip = coef[ti * 5 + 3]; k.recipm1 = coef[ti * 5 + 4];

    float mo[4][8], xt[4][8], nz[4][8];
    const int64_t base = n * 8 * S + s;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float4 a = ld_stream_f4(x_t + base + b * S);
        float4 z = ld_stream_f4(noise + base + b * S);
        xt[0][b] = a.x; xt[1][b] = a.y; xt[2][b] = a.z; xt[3][b] = a.w;
        nz[0][b] = z.x; nz[1][b] = z.y; nz[2][b] = z.z; nz[3][b] = z.w;
    }
    if (kModelOutCl) {
        const __nv_bfloat16* m = reinterpret_cast<const __nv_bfloat16*>(model_out) + (n * S + s) * mo_ld;
#pragma unroll
        for (int q = 0; q < 4; ++q) unpack8(ld_stream_u4(m + q * mo_ld), mo[q]);
    } else {
        const float* m = reinterpret_cast<const float*>(model_out) + base;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            float4 a = ld_stream_f4(m + b * S);
            mo[0][b] = a.x; mo[1][b] = a.y; mo[2][b] = a.z; mo[3][b] = a.w;
        }
    }
    float xp[4][8], pr[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) step_voxel(mo[q], xt[q], nz[q], k, clip, predict_xstart, xp[q], pr[q]);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        st_stream_f4(x_prev + base + b * S, make_float4(xp[0][b], xp[1][b], xp[2][b], xp[3][b]));
        if (pred_xstart != nullptr)
            st_stream_f4(pred_xstart + base + b * S, make_float4(pr[0][b], pr[1][b], pr[2][b], pr[3][b]));
    }
    if (x_prev_cl != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(x_prev_cl + (n * S + s + q) * xp_ld) = pack8(xp[q]);
    }
}

// scalar tail-safe variant for S % 4 != 0 or unaligned pointers (one thread per voxel)
template <bool kModelOutCl>
__global__ void __launch_bounds__(256) p_sample_step_scalar(const void* __restrict__ model_out, int64_t mo_ld,
                                                            const float* x_t,
                                                            const float* __restrict__ noise, float* x_prev,
                                                            float* __restrict__ pred_xstart,
                                                            __nv_bfloat16* __restrict__ x_prev_cl, int64_t xp_ld,
                                                            const float* __restrict__ coef,
                                                            const int64_t* __restrict__ t, int64_t T, int64_t N,
                                                            int64_t S, int clip, int predict_xstart) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    const int64_t n = idx / S, s = idx % S;
    int64_t ti = t[n];
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
    StepCoef k;
    k.c1 = coef[ti * 5 + 0]; k.c2 = coef[ti * 5 + 1]; k.sigma = coef[ti * 5 + 2];
    k.recip = coef[ti * 5 + 3]; k.recipm1 = coef[ti * 5 + 4];
    float mo[8], xt[8], nz[8], xp[8], pr[8];
    const int64_t base = n * 8 * S + s;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        xt[b] = x_t[base + b * S];
        nz[b] = noise[base + b * S];
        mo[b] = kModelOutCl
                    ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(model_out)[(n * S + s) * mo_ld + b])
                    : reinterpret_cast<const float*>(model_out)[base + b * S];
    }
    step_voxel(mo, xt, nz, k, clip, predict_xstart, xp, pr);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        x_prev[base + b * S] = xp[b];
        if (pred_xstart != nullptr) pred_xstart[base + b * S] = pr[b];
        if (x_prev_cl != nullptr) x_prev_cl[(n * S + s) * xp_ld + b] = __float2bfloat16_rn(xp[b]);
    }
}

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ nz,
                                                       float* __restrict__ out, const float* __restrict__ coef,
                                                       const int64_t* __restrict__ t, int64_t T, int64_t N,
                                                       int64_t per) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * per) return;
    int64_t ti = t[idx / per];
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
    out[idx] = coef[ti * 2] * x0[idx] + coef[ti * 2 + 1] * nz[idx];   // gaussian_diffusion.py:238-241
}

// scripts/sample.py:113-125: one thread per latent voxel writes its 2x2x2 image brick.
__global__ void __launch_bounds__(256) sample_to_image_kernel(const float* __restrict__ sample,
                                                              const float* __restrict__ cond1,
                                                              float* __restrict__ image, int64_t N, int64_t d,
                                                              int64_t h, int64_t w) {
    pdl_prologue();
    const int64_t S = d * h * w;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    const int64_t n = idx / S;
    int64_t s = idx % S;
    const int64_t ww = s % w; s /= w;
    const int64_t hh = s % h;
    const int64_t dd = s / h;
    float b[8], img[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = sample[(n * 8 + k) * S + (dd * h + hh) * w + ww];
    b[0] *= 3.0f;
    haar_synthesis(b, img);
    const int64_t H = 2 * h, W = 2 * w;
    const int64_t o = n * 8 * S + ((2 * dd) * H + 2 * hh) * W + 2 * ww;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t off = o + (i * H + j) * W;
            float v0 = img[i * 4 + j * 2], v1 = img[i * 4 + j * 2 + 1];
            v0 = (v0 <= 0.f) ? 0.f : ((v0 >= 1.f) ? 1.f : v0);   // sample[sample <= 0] = 0; sample[sample >= 1] = 1
            v1 = (v1 <= 0.f) ? 0.f : ((v1 >= 1.f) ? 1.f : v1);
            if (cond1 != nullptr) {
                float2 c = *reinterpret_cast<const float2*>(cond1 + off);
                if (c.x == 0.f) v0 = 0.f;                        // sample[cond_1 == 0] = 0
                if (c.y == 0.f) v1 = 0.f;
            }
            *reinterpret_cast<float2*>(image + off) = make_float2(v0, v1);
        }
}

// planar f32 (N,C,S) -> cl bf16 (N,S,ld): tile transpose through shared memory
__global__ void __launch_bounds__(256) planar_to_cl_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                           int64_t dst_ld, int64_t C, int64_t S) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const int64_t n = blockIdx.z;
    const int64_t s0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        int64_t c = c0 + r, s = s0 + tx;
        tile[r][tx] = (c < C && s < S) ? src[(n * C + c) * S + s] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        int64_t s = s0 + r, c = c0 + tx;
        if (s < S && c < C) dst[(n * S + s) * dst_ld + c] = __float2bfloat16_rn(tile[tx][r]);
    }
}

__global__ void __launch_bounds__(256) cl_to_planar_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld,
                                                           float* __restrict__ dst, int64_t C, int64_t S) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const int64_t n = blockIdx.z;
    const int64_t s0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        int64_t s = s0 + r, c = c0 + tx;
        tile[r][tx] = (c < C && s < S) ? __bfloat162float(src[(n * S + s) * src_ld + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        int64_t c = c0 + r, s = s0 + tx;
        if (c < C && s < S) dst[(n * C + c) * S + s] = tile[tx][r];
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_p_sample_step(const void* model_out, int64_t mo_cl_ld, const float* x_t, const float* noise,
                                   float* x_prev, float* pred_xstart, void* x_prev_cl, int64_t xp_cl_ld,
                                   const float* coef, const int64_t* t, int64_t T, int64_t N, int64_t d, int64_t h,
                                   int64_t w, int clip_denoised, int predict_xstart, void* stream) {
    FCWDM_REQUIRE(model_out && x_t && noise && x_prev && coef && t, FCWDM_ERR_INVALID,
                  "fcwdm_p_sample_step: null pointer");
    FCWDM_REQUIRE(N >= 0 && d >= 0 && h >= 0 && w >= 0 && T > 0 && mo_cl_ld >= 0, FCWDM_ERR_INVALID,
                  "fcwdm_p_sample_step: bad dimension");
    FCWDM_REQUIRE(mo_cl_ld == 0 || mo_cl_ld >= 8, FCWDM_ERR_INVALID, "fcwdm_p_sample_step: mo_cl_ld must be 0 or >= 8");
    FCWDM_REQUIRE(x_prev_cl == nullptr || xp_cl_ld >= 8, FCWDM_ERR_INVALID, "fcwdm_p_sample_step: xp_cl_ld must be >= 8");
    const int64_t S = d * h * w;
    if (N * S == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool cl = mo_cl_ld != 0;
    const bool vec = (S % 4 == 0) && aligned16(x_t) && aligned16(noise) && aligned16(x_prev) &&
                     aligned16(pred_xstart) && aligned16(model_out) && (!cl || mo_cl_ld % 8 == 0) &&
                     (x_prev_cl == nullptr || (aligned16(x_prev_cl) && xp_cl_ld % 8 == 0));
    if (vec) {
        const int64_t total = N * (S / 4);
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (cl)
            launch_k(p_sample_step_kernel<true>, dim3(grid), dim3(256), 0, st, model_out, mo_cl_ld, x_t, noise, x_prev, pred_xstart,
                                                            (__nv_bfloat16*)x_prev_cl, xp_cl_ld, coef, t, T, N, S,
                                                            clip_denoised, predict_xstart);
        else
            launch_k(p_sample_step_kernel<false>, dim3(grid), dim3(256), 0, st, model_out, 0, x_t, noise, x_prev, pred_xstart,
                                                             (__nv_bfloat16*)x_prev_cl, xp_cl_ld, coef, t, T, N, S,
                                                             clip_denoised, predict_xstart);
    } else {
        const int64_t total = N * S;
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (cl)
            launch_k(p_sample_step_scalar<true>, dim3(grid), dim3(256), 0, st, model_out, mo_cl_ld, x_t, noise, x_prev, pred_xstart,
                                                            (__nv_bfloat16*)x_prev_cl, xp_cl_ld, coef, t, T, N, S,
                                                            clip_denoised, predict_xstart);
        else
            launch_k(p_sample_step_scalar<false>, dim3(grid), dim3(256), 0, st, model_out, 0, x_t, noise, x_prev, pred_xstart,
                                                             (__nv_bfloat16*)x_prev_cl, xp_cl_ld, coef, t, T, N, S,
                                                             clip_denoised, predict_xstart);
    }
    FCWDM_CHECK_LAUNCH("fcwdm_p_sample_step");
    return FCWDM_OK;
}

extern "C" int fcwdm_q_sample(const float* x_start, const float* noise, float* out, const float* coef,
                              const int64_t* t, int64_t T, int64_t N, int64_t per_sample, void* stream) {
    FCWDM_REQUIRE(x_start && noise && out && coef && t, FCWDM_ERR_INVALID, "fcwdm_q_sample: null pointer");
    FCWDM_REQUIRE(N >= 0 && per_sample >= 0 && T > 0, FCWDM_ERR_INVALID, "fcwdm_q_sample: bad dimension");
    const int64_t total = N * per_sample;
    if (total == 0) return FCWDM_OK;
    launch_k(q_sample_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x_start, noise, out, coef, t, T,
                                                                                      N, per_sample);
    FCWDM_CHECK_LAUNCH("fcwdm_q_sample");
    return FCWDM_OK;
}

extern "C" int fcwdm_sample_to_image(const float* sample, const float* cond_1, float* image, int64_t N, int64_t d,
                                     int64_t h, int64_t w, void* stream) {
    FCWDM_REQUIRE(sample && image, FCWDM_ERR_INVALID, "fcwdm_sample_to_image: null pointer");
    FCWDM_REQUIRE(N >= 0 && d >= 0 && h >= 0 && w >= 0, FCWDM_ERR_INVALID, "fcwdm_sample_to_image: bad dimension");
    FCWDM_REQUIRE((reinterpret_cast<uintptr_t>(image) & 7) == 0 && (reinterpret_cast<uintptr_t>(cond_1) & 7) == 0,
                  FCWDM_ERR_INVALID, "fcwdm_sample_to_image: image / cond_1 must be 8-byte aligned");
    const int64_t total = N * d * h * w;
    if (total == 0) return FCWDM_OK;
    launch_k(sample_to_image_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, sample, cond_1, image, N,
                                                                                             d, h, w);
    FCWDM_CHECK_LAUNCH("fcwdm_sample_to_image");
    return FCWDM_OK;
}

extern "C" int fcwdm_planar_to_cl(const float* src, void* dst, int64_t dst_ld, int64_t N, int64_t C, int64_t S,
                                  void* stream) {
    FCWDM_REQUIRE(src && dst, FCWDM_ERR_INVALID, "fcwdm_planar_to_cl: null pointer");
    FCWDM_REQUIRE(N >= 0 && C >= 0 && S >= 0 && dst_ld >= C && N <= 65535, FCWDM_ERR_INVALID,
                  "fcwdm_planar_to_cl: bad dimension");
    if (N * C * S == 0) return FCWDM_OK;
    dim3 grid((unsigned)((S + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
    launch_k(planar_to_cl_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, dst_ld, C, S);
    FCWDM_CHECK_LAUNCH("fcwdm_planar_to_cl");
    return FCWDM_OK;
}

extern "C" int fcwdm_cl_to_planar(const void* src, int64_t src_ld, float* dst, int64_t N, int64_t C, int64_t S,
                                  void* stream) {
    FCWDM_REQUIRE(src && dst, FCWDM_ERR_INVALID, "fcwdm_cl_to_planar: null pointer");
    FCWDM_REQUIRE(N >= 0 && C >= 0 && S >= 0 && src_ld >= C && N <= 65535, FCWDM_ERR_INVALID,
                  "fcwdm_cl_to_planar: bad dimension");
    if (N * C * S == 0) return FCWDM_OK;
    dim3 grid((unsigned)((S + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
    launch_k(cl_to_planar_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, src_ld, dst, C, S);
    FCWDM_CHECK_LAUNCH("fcwdm_cl_to_planar");
    return FCWDM_OK;
}
