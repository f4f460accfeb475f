// Backward kernels of the training path (scripts/train.py -> TrainLoop.forward_backward, train_util.py:396-460 ->
// autograd through WavUNetModel): GroupNorm+SiLU backward, per-channel column sums (bias / timestep-embedding
// gradients), channels-last Haar DWT/IDWT adjoints, the small dense layers' backward, and a fused AdamW step.
// All HBM-bound elementwise / reduction passes on channels-last bf16 gradients with fp32 math; parameter
// gradients are fp32 and ACCUMULATE into their destination (the caller zeroes them once per step), which is what
// weight-tied ResBlocks (wunet.py:647-673) need.
#include <cstdlib>
#include "common.cuh"

namespace fcwdm {

constexpr int kTrThreads = 256;
constexpr int kTrReplicas = FCWDM_GN_STAT_REPLICAS;

// sigmoid through the hardware tanh (one MUFU op, relative error ~2^-11: far below the bf16 rounding of the gradients)
__device__ __forceinline__ float sigmoid_f(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}

// per-channel (scale, shift) of x_hat = x * scale + shift from the (sum, sumsq) statistics of fcwdm_groupnorm_stats
__device__ __forceinline__ void gn_channel_norm(const double* __restrict__ stats, int n, int g, int G, double cnt,
                                                float eps, float& rstd, float& mean) {
    double sum = 0.0, sq = 0.0;
    for (int r = 0; r < kTrReplicas; ++r) {
        sum += stats[(((int64_t)n * kTrReplicas + r) * G + g) * 2 + 0];
        sq += stats[(((int64_t)n * kTrReplicas + r) * G + g) * 2 + 1];
    }
    const double m = sum / cnt;
    double var = sq / cnt - m * m;
    var = var < 0.0 ? 0.0 : var;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    mean = (float)m;
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm(+SiLU) backward, pass 1: sums[n][replica][c] = (sum_v dz, sum_v dz * x_hat),
// dz = dy * silu'(gamma * x_hat + beta) (or dy without SiLU)
// ---------------------------------------------------------------------------------------------------
template <bool kSilu>
__global__ void __launch_bounds__(kTrThreads) gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld,
                                                                   const __nv_bfloat16* __restrict__ dy, int64_t dy_ld,
                                                                   const double* __restrict__ stats,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   double* __restrict__ sums, int64_t S, int C, int G,
                                                                   float eps) {
    pdl_prologue();
    extern __shared__ float sm[];   // [16][kTrThreads] partials, then [4][C]: xs, xh, gam, bet; reused as [2][C] sums
    float* part = sm;
    float* coef = sm + 16 * kTrThreads;
    const int n = blockIdx.y;
    const int cpg = C / G;
    const double cnt = (double)S * (double)cpg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float rstd, mean;
        gn_channel_norm(stats, n, c / cpg, G, cnt, eps, rstd, mean);
        coef[c] = rstd;
        coef[C + c] = -mean * rstd;
        coef[2 * C + c] = gamma[c];
        coef[3 * C + c] = beta[c];
    }
    __syncthreads();
    const int C8 = C >> 3;
    const int chunk = threadIdx.x % C8;
    const int lane = threadIdx.x / C8;
    const int vpb = blockDim.x / C8;
    float xs[8], xh[8], gm[8], bt[8], a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        xs[j] = coef[chunk * 8 + j];
        xh[j] = coef[C + chunk * 8 + j];
        gm[j] = coef[2 * C + chunk * 8 + j];
        bt[j] = coef[3 * C + chunk * 8 + j];
        a[j] = b[j] = 0.f;
    }
    const __nv_bfloat16* xb = x + (int64_t)n * S * x_ld + chunk * 8;
    const __nv_bfloat16* db = dy + (int64_t)n * S * dy_ld + chunk * 8;
    const int64_t stride = (int64_t)gridDim.x * vpb;
    auto accumulate = [&](const uint4& ux, const uint4& ud) {
        float fx[8], fd[8];
        unpack8(ux, fx);
        unpack8(ud, fd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = fmaf(fx[j], xs[j], xh[j]);
            float dz = fd[j];
            if (kSilu) {
                const float z = fmaf(h, gm[j], bt[j]);
                const float sg = sigmoid_f(z);
                dz *= sg * fmaf(z, 1.0f - sg, 1.0f);
            }
            a[j] += dz;
            b[j] = fmaf(dz, h, b[j]);
        }
    };
    int64_t v = (int64_t)blockIdx.x * vpb + lane;
    // 8 independent 128-bit loads in flight per thread (4 voxels of x and dy), then the arithmetic
    for (; v + 3 * stride < S; v += 4 * stride) {
        uint4 ux[4], ud[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ux[k] = ld_stream_u4(xb + (v + k * stride) * x_ld);
            ud[k] = ld_stream_u4(db + (v + k * stride) * dy_ld);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) accumulate(ux[k], ud[k]);
    }
    for (; v < S; v += stride) accumulate(ld_stream_u4(xb + v * x_ld), ld_stream_u4(db + v * dy_ld));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        part[j * kTrThreads + threadIdx.x] = a[j];
        part[(8 + j) * kTrThreads + threadIdx.x] = b[j];
    }
    __syncthreads();
    double* dst = sums + (((int64_t)n * kTrReplicas + (blockIdx.x % kTrReplicas)) * C) * 2;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        const int comp = i / C, c = i % C;
        const float* p = part + (comp * 8 + (c & 7)) * kTrThreads + (c >> 3);
        float acc = 0.f;
        for (int l = 0; l < vpb; ++l) acc += p[l * C8];
        atomicAdd(&dst[c * 2 + comp], (double)acc);
    }
}

// pass 2: dx = rstd * (gamma * dz - (S1 + x_hat * S2) / M) (+ acc), S1 = sum_{c in group} gamma_c A_c,
// S2 = sum gamma_c B_c, M = S * channels_per_group; block (0, n) also adds A, B into dbeta, dgamma.
// kColsum: also cs[n][c] += sum_v dx[n,v,c] over the STORED (bf16) values -- the bias / timestep-embedding gradient of the
// conv that produced x, which otherwise costs fcwdm_colsum_cl one more pass over dx.
template <bool kSilu, bool kColsum>
__global__ void __launch_bounds__(kTrThreads, 2) gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld,
                                                                  const __nv_bfloat16* __restrict__ dy, int64_t dy_ld,
                                                                  const double* __restrict__ stats,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const double* __restrict__ sums,
                                                                  const __nv_bfloat16* __restrict__ acc, int64_t acc_ld,
                                                                  __nv_bfloat16* __restrict__ dx, int64_t dx_ld,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                  float* __restrict__ cs, int64_t cs_ld,
                                                                  int64_t S, int C, int G, float eps) {
    pdl_prologue();
    extern __shared__ float sm[];   // [7][C]: xs, xh, gam, bet, k2, k3, scratch(gamma*A) ; scratch2 (gamma*B) at [7]
                                    // kColsum: reused as [8][blockDim] partial column sums at the end
    const int n = blockIdx.y;
    const int cpg = C / G;
    const double cnt = (double)S * (double)cpg;
    float* ga = sm + 6 * C;
    float* gb = sm + 7 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float rstd, mean;
        gn_channel_norm(stats, n, c / cpg, G, cnt, eps, rstd, mean);
        double A = 0.0, B = 0.0;
        for (int r = 0; r < kTrReplicas; ++r) {
            A += sums[(((int64_t)n * kTrReplicas + r) * C + c) * 2 + 0];
            B += sums[(((int64_t)n * kTrReplicas + r) * C + c) * 2 + 1];
        }
        sm[c] = rstd;
        sm[C + c] = -mean * rstd;
        sm[2 * C + c] = gamma[c];
        sm[3 * C + c] = beta[c];
        ga[c] = gamma[c] * (float)A;
        gb[c] = gamma[c] * (float)B;
        if (blockIdx.x == 0) {
            if (dbeta != nullptr) atomicAdd(dbeta + c, (float)A);
            if (dgamma != nullptr) atomicAdd(dgamma + c, (float)B);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g0 = (c / cpg) * cpg;
        float s1 = 0.f, s2 = 0.f;
        for (int e = 0; e < cpg; ++e) {
            s1 += ga[g0 + e];
            s2 += gb[g0 + e];
        }
        const float inv = (float)(1.0 / cnt);
        sm[4 * C + c] = sm[c] * s2 * inv;    // k2
        sm[5 * C + c] = sm[c] * s1 * inv;    // k3
    }
    __syncthreads();
    const int C8 = C >> 3;
    const int chunk = threadIdx.x % C8;
    const int lane = threadIdx.x / C8;
    const int vpb = blockDim.x / C8;
    float xs[8], xh[8], gm[8], bt[8], k1[8], k2[8], k3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = chunk * 8 + j;
        xs[j] = sm[c];
        xh[j] = sm[C + c];
        gm[j] = sm[2 * C + c];
        bt[j] = sm[3 * C + c];
        k1[j] = xs[j] * gm[j];
        k2[j] = sm[4 * C + c];
        k3[j] = sm[5 * C + c];
    }
    const __nv_bfloat16* xb = x + (int64_t)n * S * x_ld + chunk * 8;
    const __nv_bfloat16* db = dy + (int64_t)n * S * dy_ld + chunk * 8;
    const __nv_bfloat16* ab = acc ? acc + (int64_t)n * S * acc_ld + chunk * 8 : nullptr;
    __nv_bfloat16* ob = dx + (int64_t)n * S * dx_ld + chunk * 8;
    const int64_t stride = (int64_t)gridDim.x * vpb;
    float csum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[j] = 0.f;
    auto apply = [&](const uint4& ux, const uint4& ud, const uint4& ua, int64_t v) {
        float fx[8], fd[8], fa[8];
        unpack8(ux, fx);
        unpack8(ud, fd);
        if (ab != nullptr) unpack8(ua, fa);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = fmaf(fx[j], xs[j], xh[j]);
            float dz = fd[j];
            if (kSilu) {
                const float z = fmaf(h, gm[j], bt[j]);
                const float sg = sigmoid_f(z);
                dz *= sg * fmaf(z, 1.0f - sg, 1.0f);
            }
            float o = k1[j] * dz - h * k2[j] - k3[j];
            if (ab != nullptr) o += fa[j];
            fd[j] = o;
        }
        const uint4 packed = pack8(fd);
        st_stream_u4(ob + v * dx_ld, packed);
        if (kColsum) {
            unpack8(packed, fd);
#pragma unroll
            for (int j = 0; j < 8; ++j) csum[j] += fd[j];
        }
    };
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    int64_t v = (int64_t)blockIdx.x * vpb + lane;
    for (; v + 3 * stride < S; v += 4 * stride) {          // up to 12 independent 128-bit loads in flight per thread
        uint4 ux[4], ud[4], ua[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ux[k] = ld_stream_u4(xb + (v + k * stride) * x_ld);
            ud[k] = ld_stream_u4(db + (v + k * stride) * dy_ld);
            ua[k] = ab != nullptr ? ld_stream_u4(ab + (v + k * stride) * acc_ld) : zero4;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) apply(ux[k], ud[k], ua[k], v + k * stride);
    }
    for (; v < S; v += stride)
        apply(ld_stream_u4(xb + v * x_ld), ld_stream_u4(db + v * dy_ld),
              ab != nullptr ? ld_stream_u4(ab + v * acc_ld) : zero4, v);
    if (kColsum) {
        __syncthreads();                                   // the coefficient table is dead: reuse it for the partials
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[j * kTrThreads + threadIdx.x] = csum[j];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float* p = sm + (c & 7) * kTrThreads + (c >> 3);
            float t = 0.f;
            for (int l = 0; l < vpb; ++l) t += p[l * C8];
            atomicAdd(cs + (int64_t)n * cs_ld + c, t);
        }
    }
}

// out_sample[n][c] += part[n][c]; out_total[c] += sum_n part[n][c]: hands the column sums gn_bwd_apply_kernel<.., true>
// left in a scratch buffer to their two destinations (timestep-embedding gradient per sample, conv bias gradient)
__global__ void __launch_bounds__(256) colsum_scatter_kernel(const float* __restrict__ part, int64_t p_ld,
                                                             float* __restrict__ out_sample, int64_t os_ld,
                                                             float* __restrict__ out_total, int N, int C) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float t = 0.f;
    for (int n = 0; n < N; ++n) {
        const float v = part[(int64_t)n * p_ld + c];
        if (out_sample != nullptr) out_sample[(int64_t)n * os_ld + c] += v;
        t += v;
    }
    if (out_total != nullptr) out_total[c] += t;
}

// ---------------------------------------------------------------------------------------------------
// column sums of a channels-last bf16 tensor: out_sample[n][c] += sum_v x[n,v,c]; out_total[c] += sum_{n,v}
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTrThreads) colsum_cl_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld,
                                                               float* __restrict__ out_sample, int64_t os_ld,
                                                               float* __restrict__ out_total, int64_t S, int C) {
    pdl_prologue();
    extern __shared__ float sm[];   // [8][kTrThreads]
    const int n = blockIdx.y;
    const int C8 = C >> 3;
    const int chunk = threadIdx.x % C8;
    const int lane = threadIdx.x / C8;
    const int vpb = blockDim.x / C8;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    const __nv_bfloat16* base = x + (int64_t)n * S * ld + chunk * 8;
    const int64_t stride = (int64_t)gridDim.x * vpb;
    int64_t v = (int64_t)blockIdx.x * vpb + lane;
    for (; v + 3 * stride < S; v += 4 * stride) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = ld_stream_u4(base + (v + k * stride) * ld);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += f[j];
        }
    }
    for (; v < S; v += stride) {
        float f[8];
        unpack8(ld_stream_u4(base + v * ld), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[j * kTrThreads + threadIdx.x] = a[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float* p = sm + (c & 7) * kTrThreads + (c >> 3);
        float acc = 0.f;
        for (int l = 0; l < vpb; ++l) acc += p[l * C8];
        if (out_sample != nullptr) atomicAdd(out_sample + (int64_t)n * os_ld + c, acc);
        if (out_total != nullptr) atomicAdd(out_total + c, acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// adjoints of the channels-last Haar kernels (fcwdm_dwt3d_cl / fcwdm_idwt3d_cl).  The Haar analysis matrix is
// orthonormal, so the adjoint of the analysis is the synthesis and vice versa (DWT_IDWT_Functions.py:139-156,
// 184-208).
// ---------------------------------------------------------------------------------------------------
// dx = IDWT(lll_scale * dlll, hi_scale * dhi [0 if dhi == null]) (+ acc)
__global__ void __launch_bounds__(256) dwt3d_cl_bwd_kernel(const __nv_bfloat16* __restrict__ dlll, int64_t lll_ld,
                                                           const __nv_bfloat16* __restrict__ dhi, int64_t hi_ld,
                                                           int64_t hi_sb, const __nv_bfloat16* __restrict__ acc,
                                                           int64_t acc_ld, __nv_bfloat16* __restrict__ dx, int64_t dx_ld,
                                                           int64_t total, int64_t D, int64_t H, int64_t W, int64_t C,
                                                           float lll_scale, float hi_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t C8 = C >> 3, D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    const int64_t cq = idx % C8;
    int64_t t = idx / C8;
    const int64_t ww = t % W2; t /= W2;
    const int64_t hh = t % H2; t /= H2;
    const int64_t dd = t % D2;
    const int64_t n = t / D2;
    const int64_t vox = ((n * D2 + dd) * H2 + hh) * W2 + ww;
    float bnd[8][8];
    unpack8(ld_stream_u4(dlll + vox * lll_ld + cq * 8), bnd[0]);
#pragma unroll
    for (int c = 0; c < 8; ++c) bnd[0][c] *= lll_scale;
    if (dhi != nullptr) {
#pragma unroll
        for (int b = 1; b < 8; ++b) {
            unpack8(ld_stream_u4(dhi + (b - 1) * hi_sb + vox * hi_ld + cq * 8), bnd[b]);
#pragma unroll
            for (int c = 0; c < 8; ++c) bnd[b][c] *= hi_scale;
        }
    } else {
#pragma unroll
        for (int b = 1; b < 8; ++b)
#pragma unroll
            for (int c = 0; c < 8; ++c) bnd[b][c] = 0.f;
    }
    float out[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float b[8], xo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] = bnd[q][c];
        haar_synthesis(b, xo);
#pragma unroll
        for (int q = 0; q < 8; ++q) out[q][c] = xo[q];
    }
    const int64_t off0 = ((n * D + 2 * dd) * H + 2 * hh) * W + 2 * ww;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int64_t o = off0 + (i * H + j) * W + k;
                if (acc != nullptr) {
                    float fa[8];
                    unpack8(ld_stream_u4(acc + o * acc_ld + cq * 8), fa);
#pragma unroll
                    for (int c = 0; c < 8; ++c) out[i * 4 + j * 2 + k][c] += fa[c];
                }
                *reinterpret_cast<uint4*>(dx + o * dx_ld + cq * 8) = pack8(out[i * 4 + j * 2 + k]);
            }
}

// (dlll, dhi) = DWT(dy): dlll = lll_scale * LLL (+ lll_acc); dhi (=, or += when hi_accumulate) the 7 high bands
__global__ void __launch_bounds__(256) idwt3d_cl_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_ld,
                                                            const __nv_bfloat16* __restrict__ lll_acc, int64_t acc_ld,
                                                            __nv_bfloat16* __restrict__ dlll, int64_t lll_ld,
                                                            __nv_bfloat16* __restrict__ dhi, int64_t hi_ld, int64_t hi_sb,
                                                            int hi_accumulate, int64_t total, int64_t D, int64_t H,
                                                            int64_t W, int64_t C, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t C8 = C >> 3, D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    const int64_t cq = idx % C8;
    int64_t t = idx / C8;
    const int64_t ww = t % W2; t /= W2;
    const int64_t hh = t % H2; t /= H2;
    const int64_t dd = t % D2;
    const int64_t n = t / D2;
    const __nv_bfloat16* src = dy + (((n * D + 2 * dd) * H + 2 * hh) * W + 2 * ww) * dy_ld + cq * 8;
    float in[8][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                unpack8(ld_stream_u4(src + ((i * H + j) * W + k) * dy_ld), in[i * 4 + j * 2 + k]);
    float ob[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float xin[8], b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) xin[q] = in[q][c];
        haar_analysis(xin, b);
#pragma unroll
        for (int q = 0; q < 8; ++q) ob[q][c] = b[q];
    }
    const int64_t vox = ((n * D2 + dd) * H2 + hh) * W2 + ww;
    if (dlll != nullptr) {
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = ob[0][c] * lll_scale;
        if (lll_acc != nullptr) {
            float fa[8];
            unpack8(ld_stream_u4(lll_acc + vox * acc_ld + cq * 8), fa);
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] += fa[c];
        }
        *reinterpret_cast<uint4*>(dlll + vox * lll_ld + cq * 8) = pack8(o);
    }
    if (dhi != nullptr) {
#pragma unroll
        for (int b = 1; b < 8; ++b) {
            __nv_bfloat16* p = dhi + (b - 1) * hi_sb + vox * hi_ld + cq * 8;
            if (hi_accumulate) {
                float fa[8];
                unpack8(*reinterpret_cast<const uint4*>(p), fa);
#pragma unroll
                for (int c = 0; c < 8; ++c) ob[b][c] += fa[c];
            }
            *reinterpret_cast<uint4*>(p) = pack8(ob[b]);
        }
    }
}

// y = a + b on channels-last bf16 buffers (gradient fan-in where no producing kernel can absorb the add)
__global__ void __launch_bounds__(256) add_cl_kernel(const __nv_bfloat16* __restrict__ a, int64_t a_ld,
                                                     const __nv_bfloat16* __restrict__ b, int64_t b_ld,
                                                     __nv_bfloat16* __restrict__ y, int64_t y_ld, int64_t rows, int C8) {
    pdl_prologue();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * C8) return;
    const int64_t r = idx / C8;
    const int c = (int)(idx % C8) * 8;
    float fa[8], fb[8];
    unpack8(ld_stream_u4(a + r * a_ld + c), fa);
    unpack8(ld_stream_u4(b + r * b_ld + c), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    *reinterpret_cast<uint4*>(y + r * y_ld + c) = pack8(fa);
}

// ---------------------------------------------------------------------------------------------------
// dense layers' backward: y[n][m] = b[m] + sum_k act(x[n][k]) W[m][k]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_f(float v, int kind) { return kind == 1 ? v / (1.0f + expf(-v)) : v; }
__device__ __forceinline__ float act_grad(float v, int kind) {
    if (kind != 1) return 1.0f;
    const float s = 1.0f / (1.0f + expf(-v));
    return s * (1.0f + v * (1.0f - s));
}

__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           int64_t dy_ld, float* __restrict__ dW, float* __restrict__ db,
                                                           int64_t N, int64_t K, int64_t M, int act_in) {
    pdl_prologue();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * K) return;
    const int64_t m = idx / K, k = idx % K;
    float acc = 0.f, accb = 0.f;
    for (int64_t n = 0; n < N; ++n) {
        const float g = dy[n * dy_ld + m];
        acc = fmaf(g, act_f(x[n * K + k], act_in), acc);
        accb += g;
    }
    dW[idx] += acc;
    if (k == 0 && db != nullptr) db[m] += accb;
}

// one block per (n, 32 input features, slice of the M outputs): 8 thread groups stride over the slice (coalesced along
// k), tree-add, then ONE atomic per feature into dx (zeroed by the wrapper unless accumulating); the activation
// derivative distributes over the partial sums
constexpr int kLinBwdSlices = 16;
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ dy, int64_t dy_ld,
                                                           float* __restrict__ dx, int64_t N, int64_t K, int64_t M,
                                                           int act_in) {
    pdl_prologue();
    __shared__ float part[8][32];
    const int c = threadIdx.x & 31, lane = threadIdx.x >> 5;
    const int64_t n = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * 32 + c;
    const int64_t per = (M + gridDim.z - 1) / gridDim.z;
    const int64_t m0 = (int64_t)blockIdx.z * per, m1 = m0 + per < M ? m0 + per : M;
    float acc = 0.f;
    if (k < K)
        for (int64_t m = m0 + lane; m < m1; m += 8) acc = fmaf(dy[n * dy_ld + m], W[m * K + k], acc);
    part[lane][c] = acc;
    __syncthreads();
    if (lane == 0 && k < K) {
        float v = ((part[0][c] + part[1][c]) + (part[2][c] + part[3][c])) + ((part[4][c] + part[5][c]) + (part[6][c] + part[7][c]));
        atomicAdd(dx + n * K + k, v * act_grad(x[n * K + k], act_in));
    }
}

// ---------------------------------------------------------------------------------------------------
// fused AdamW over a flat fp32 parameter range (torch.optim.AdamW semantics, train_util.py:75-82,391)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                    float beta1, float beta2, float eps, float weight_decay,
                                                    float bias1, float bias2_sqrt, float grad_scale) {
    pdl_prologue();
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 3 < n) {
        float4 pp = *reinterpret_cast<float4*>(p + i4);
        const float4 gg = *reinterpret_cast<const float4*>(g + i4);
        float4 mm = *reinterpret_cast<float4*>(m + i4);
        float4 vv = *reinterpret_cast<float4*>(v + i4);
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * grad_scale;
            pa[j] *= 1.0f - lr * weight_decay;
            ma[j] = beta1 * ma[j] + (1.0f - beta1) * gr;
            va[j] = beta2 * va[j] + (1.0f - beta2) * gr * gr;
            pa[j] -= (lr / bias1) * ma[j] / (sqrtf(va[j]) / bias2_sqrt + eps);
        }
        *reinterpret_cast<float4*>(p + i4) = pp;
        *reinterpret_cast<float4*>(m + i4) = mm;
        *reinterpret_cast<float4*>(v + i4) = vv;
    } else {
        for (int64_t i = i4; i < n; ++i) {
            const float gr = g[i] * grad_scale;
            float pv = p[i] * (1.0f - lr * weight_decay);
            const float mv = beta1 * m[i] + (1.0f - beta1) * gr;
            const float vv = beta2 * v[i] + (1.0f - beta2) * gr * gr;
            pv -= (lr / bias1) * mv / (sqrtf(vv) / bias2_sqrt + eps);
            p[i] = pv; m[i] = mv; v[i] = vv;
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

static int gn_bwd_check(const char* fn, int64_t N, int64_t S, int64_t C, int64_t G) {
    FCWDM_REQUIRE(N >= 0 && S >= 0 && C > 0 && G > 0 && N <= 65535, FCWDM_ERR_INVALID, "%s: bad dimension", fn);
    FCWDM_REQUIRE(C % G == 0, FCWDM_ERR_INVALID, "%s: C (%lld) not divisible by G (%lld)", fn, (long long)C, (long long)G);
    FCWDM_REQUIRE(C % 8 == 0 && C <= 1024, FCWDM_ERR_UNSUPPORTED, "%s: C must be a multiple of 8, at most 1024 (got %lld)",
                  fn, (long long)C);
    return FCWDM_OK;
}

// blocks_per_sm caps the grid: measured (tools/gnbwd_probe.py, 2 x 64 ch x 1 M voxels / 2 x 128 x 125 k) the two GroupNorm
// backward passes want ONE resident wave (2 blocks per SM: 302 -> 286 us and 104 -> 83 us against a cap of 8), the
// plain column sum wants more blocks in flight (56 us at 8 per SM, 98 us at 2).
static inline dim3 slab_grid(int64_t N, int64_t S, int64_t C, int per_thread, int blocks_per_sm) {
    const int64_t vpb = kTrThreads / (C / 8);
    int64_t blocks = (S + vpb * per_thread - 1) / (vpb * per_thread);
    static const int env_per_sm = getenv("FCWDM_TR_BLOCKS_PER_SM") ? atoi(getenv("FCWDM_TR_BLOCKS_PER_SM")) : 0;
    const int per_sm = env_per_sm > 0 ? env_per_sm : blocks_per_sm;
    const int64_t cap = (int64_t)num_sms() * per_sm / (N > 8 ? 8 : N);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, (unsigned)N);
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_groupnorm_bwd(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, const double* stats,
                                   const float* gamma, const float* beta, double* sums, const void* acc, int64_t acc_ld,
                                   void* dx, int64_t dx_ld, float* dgamma, float* dbeta, int64_t N, int64_t S, int64_t C,
                                   int64_t G, float eps, int silu, void* stream) {
    return fcwdm_groupnorm_bwd_colsum(x, x_ld, dy, dy_ld, stats, gamma, beta, sums, acc, acc_ld, dx, dx_ld, dgamma, dbeta,
                                      nullptr, 0, N, S, C, G, eps, silu, stream);
}

extern "C" int fcwdm_groupnorm_bwd_colsum(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, const double* stats,
                                          const float* gamma, const float* beta, double* sums, const void* acc,
                                          int64_t acc_ld, void* dx, int64_t dx_ld, float* dgamma, float* dbeta,
                                          float* colsum, int64_t colsum_ld, int64_t N, int64_t S, int64_t C, int64_t G,
                                          float eps, int silu, void* stream) {
    FCWDM_REQUIRE(x && dy && stats && gamma && beta && sums && dx, FCWDM_ERR_INVALID, "fcwdm_groupnorm_bwd: null pointer");
    FCWDM_REQUIRE(colsum == nullptr || colsum_ld >= C, FCWDM_ERR_INVALID, "fcwdm_groupnorm_bwd: bad column-sum stride");
    int rc = gn_bwd_check("fcwdm_groupnorm_bwd", N, S, C, G);
    if (rc) return rc;
    FCWDM_REQUIRE(x_ld >= C && dy_ld >= C && dx_ld >= C && x_ld % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0 &&
                      (acc == nullptr || (acc_ld >= C && acc_ld % 8 == 0)),
                  FCWDM_ERR_INVALID, "fcwdm_groupnorm_bwd: bad leading dimension");
    FCWDM_REQUIRE(al16(x) && al16(dy) && al16(dx) && al16(acc), FCWDM_ERR_INVALID,
                  "fcwdm_groupnorm_bwd: pointers must be 16-byte aligned");
    if (N * S == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * N * C * kTrReplicas, st);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_groupnorm_bwd: memset failed (%s)", cudaGetErrorString(e));
    const dim3 grid = slab_grid(N, S, C, 4, 2);
    const dim3 blk((unsigned)((kTrThreads / (C / 8)) * (C / 8)));      // 256, or e.g. 240 for C = 192 / 384
    const size_t sm1 = (16 * kTrThreads + 4 * C) * sizeof(float);
    const size_t sm2 = (colsum != nullptr && 8 * C < 8 * kTrThreads ? 8 * kTrThreads : 8 * C) * sizeof(float);
    const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *dp = (const __nv_bfloat16*)dy, *ap = (const __nv_bfloat16*)acc;
    if (silu)
        launch_k(gn_bwd_reduce_kernel<true>, grid, blk, sm1, st, xp, x_ld, dp, dy_ld, stats, gamma, beta, sums,
                 S, (int)C, (int)G, eps);
    else
        launch_k(gn_bwd_reduce_kernel<false>, grid, blk, sm1, st, xp, x_ld, dp, dy_ld, stats, gamma, beta, sums,
                 S, (int)C, (int)G, eps);
#define FCWDM_GN_BWD_APPLY(SILU, CS)                                                                                    \
    launch_k(gn_bwd_apply_kernel<SILU, CS>, grid, blk, sm2, st, xp, x_ld, dp, dy_ld, stats, gamma, beta,                \
             (const double*)sums, ap, acc_ld, (__nv_bfloat16*)dx, dx_ld, dgamma, dbeta, colsum, colsum_ld, S, (int)C,   \
             (int)G, eps)
    if (silu) {
        if (colsum != nullptr) FCWDM_GN_BWD_APPLY(true, true); else FCWDM_GN_BWD_APPLY(true, false);
    } else {
        if (colsum != nullptr) FCWDM_GN_BWD_APPLY(false, true); else FCWDM_GN_BWD_APPLY(false, false);
    }
#undef FCWDM_GN_BWD_APPLY
    FCWDM_CHECK_LAUNCH("fcwdm_groupnorm_bwd");
    return FCWDM_OK;
}

extern "C" int fcwdm_colsum_scatter(const float* part, int64_t part_ld, float* out_sample, int64_t os_ld, float* out_total,
                                    int64_t N, int64_t C, void* stream) {
    FCWDM_REQUIRE(part && (out_sample || out_total), FCWDM_ERR_INVALID, "fcwdm_colsum_scatter: null pointer");
    FCWDM_REQUIRE(N >= 0 && C >= 0 && part_ld >= C && (out_sample == nullptr || os_ld >= C), FCWDM_ERR_INVALID,
                  "fcwdm_colsum_scatter: bad dimension");
    if (N * C == 0) return FCWDM_OK;
    launch_k(colsum_scatter_kernel, dim3((unsigned)((C + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, part, part_ld,
             out_sample, os_ld, out_total, (int)N, (int)C);
    FCWDM_CHECK_LAUNCH("fcwdm_colsum_scatter");
    return FCWDM_OK;
}

extern "C" int fcwdm_colsum_cl(const void* x, int64_t ld, float* out_sample, int64_t os_ld, float* out_total, int64_t N,
                               int64_t S, int64_t C, void* stream) {
    FCWDM_REQUIRE(x && (out_sample || out_total), FCWDM_ERR_INVALID, "fcwdm_colsum_cl: null pointer");
    FCWDM_REQUIRE(N >= 0 && S >= 0 && C > 0 && N <= 65535, FCWDM_ERR_INVALID, "fcwdm_colsum_cl: bad dimension");
    FCWDM_REQUIRE(C % 8 == 0 && C <= 2048 && ld >= C && ld % 8 == 0 && al16(x), FCWDM_ERR_UNSUPPORTED,
                  "fcwdm_colsum_cl: C must be a multiple of 8 (<= 2048), ld >= C, 16-byte aligned");
    if (N * S == 0) return FCWDM_OK;
    launch_k(colsum_cl_kernel, slab_grid(N, S, C, 8, 8), dim3((unsigned)((kTrThreads / (C / 8)) * (C / 8))), 8 * kTrThreads * sizeof(float),
             (cudaStream_t)stream, (const __nv_bfloat16*)x, ld, out_sample, os_ld, out_total, S, (int)C);
    FCWDM_CHECK_LAUNCH("fcwdm_colsum_cl");
    return FCWDM_OK;
}

extern "C" int fcwdm_dwt3d_cl_bwd(const void* dlll, int64_t lll_ld, const void* dhi, int64_t hi_ld, int64_t hi_sb,
                                  const void* acc, int64_t acc_ld, void* dx, int64_t dx_ld, int64_t N, int64_t D,
                                  int64_t H, int64_t W, int64_t C, float lll_scale, float hi_scale, void* stream) {
    FCWDM_REQUIRE(dlll && dx, FCWDM_ERR_INVALID, "fcwdm_dwt3d_cl_bwd: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0,
                  FCWDM_ERR_INVALID, "fcwdm_dwt3d_cl_bwd: bad dimension (D, H, W even)");
    FCWDM_REQUIRE(C % 8 == 0 && lll_ld % 8 == 0 && hi_ld % 8 == 0 && hi_sb % 8 == 0 && acc_ld % 8 == 0 && dx_ld % 8 == 0 &&
                      al16(dlll) && al16(dhi) && al16(acc) && al16(dx),
                  FCWDM_ERR_UNSUPPORTED, "fcwdm_dwt3d_cl_bwd: C and strides must be multiples of 8, pointers 16-byte aligned");
    const int64_t total = N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(dwt3d_cl_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)dlll, lll_ld, (const __nv_bfloat16*)dhi, hi_ld, hi_sb, (const __nv_bfloat16*)acc, acc_ld,
             (__nv_bfloat16*)dx, dx_ld, total, D, H, W, C, lll_scale, hi_scale);
    FCWDM_CHECK_LAUNCH("fcwdm_dwt3d_cl_bwd");
    return FCWDM_OK;
}

extern "C" int fcwdm_idwt3d_cl_bwd(const void* dy, int64_t dy_ld, const void* lll_acc, int64_t acc_ld, void* dlll,
                                   int64_t lll_ld, void* dhi, int64_t hi_ld, int64_t hi_sb, int hi_accumulate, int64_t N,
                                   int64_t D, int64_t H, int64_t W, int64_t C, float lll_scale, void* stream) {
    FCWDM_REQUIRE(dy && (dlll || dhi), FCWDM_ERR_INVALID, "fcwdm_idwt3d_cl_bwd: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0,
                  FCWDM_ERR_INVALID, "fcwdm_idwt3d_cl_bwd: bad dimension (D, H, W even)");
    FCWDM_REQUIRE(C % 8 == 0 && lll_ld % 8 == 0 && hi_ld % 8 == 0 && hi_sb % 8 == 0 && acc_ld % 8 == 0 && dy_ld % 8 == 0 &&
                      al16(dlll) && al16(dhi) && al16(lll_acc) && al16(dy),
                  FCWDM_ERR_UNSUPPORTED, "fcwdm_idwt3d_cl_bwd: C and strides must be multiples of 8, pointers 16-byte aligned");
    const int64_t total = N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(idwt3d_cl_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)lll_acc, acc_ld, (__nv_bfloat16*)dlll, lll_ld,
             (__nv_bfloat16*)dhi, hi_ld, hi_sb, hi_accumulate, total, D, H, W, C, lll_scale);
    FCWDM_CHECK_LAUNCH("fcwdm_idwt3d_cl_bwd");
    return FCWDM_OK;
}

extern "C" int fcwdm_add_cl(const void* a, int64_t a_ld, const void* b, int64_t b_ld, void* y, int64_t y_ld, int64_t rows,
                            int64_t C, void* stream) {
    FCWDM_REQUIRE(a && b && y, FCWDM_ERR_INVALID, "fcwdm_add_cl: null pointer");
    FCWDM_REQUIRE(rows >= 0 && C > 0 && C % 8 == 0 && a_ld % 8 == 0 && b_ld % 8 == 0 && y_ld % 8 == 0 && al16(a) &&
                      al16(b) && al16(y),
                  FCWDM_ERR_INVALID, "fcwdm_add_cl: C and strides must be multiples of 8, pointers 16-byte aligned");
    const int64_t total = rows * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(add_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)a, a_ld, (const __nv_bfloat16*)b, b_ld, (__nv_bfloat16*)y, y_ld, rows, (int)(C / 8));
    FCWDM_CHECK_LAUNCH("fcwdm_add_cl");
    return FCWDM_OK;
}

extern "C" int fcwdm_linear_bwd(const float* x, const float* W, const float* dy, int64_t dy_ld, float* dx, float* dW,
                                float* db, int64_t N, int64_t K, int64_t M, int act_in, int accumulate_dx, void* stream) {
    FCWDM_REQUIRE(x && dy, FCWDM_ERR_INVALID, "fcwdm_linear_bwd: null pointer");
    FCWDM_REQUIRE(N >= 0 && K > 0 && M >= 0 && dy_ld >= M && (act_in == 0 || act_in == 1), FCWDM_ERR_INVALID,
                  "fcwdm_linear_bwd: bad argument");
    FCWDM_REQUIRE(dx == nullptr || W != nullptr, FCWDM_ERR_INVALID, "fcwdm_linear_bwd: dx needs W");
    if (N * M == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dW != nullptr) {
        launch_k(linear_bwd_w_kernel, dim3((unsigned)((M * K + 255) / 256)), dim3(256), 0, st, x, dy, dy_ld, dW, db, N, K, M,
                 act_in);
        FCWDM_CHECK_LAUNCH("fcwdm_linear_bwd (dW)");
    }
    if (dx != nullptr) {
        if (!accumulate_dx) {
            cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * N * K, st);
            FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_linear_bwd: memset failed (%s)", cudaGetErrorString(e));
        }
        const unsigned slices = (unsigned)(M >= 64 * kLinBwdSlices ? kLinBwdSlices : 1);
        launch_k(linear_bwd_x_kernel, dim3((unsigned)((K + 31) / 32), (unsigned)N, slices), dim3(256), 0, st, x, W, dy, dy_ld, dx,
                 N, K, M, act_in);
        FCWDM_CHECK_LAUNCH("fcwdm_linear_bwd (dx)");
    }
    return FCWDM_OK;
}

extern "C" int fcwdm_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
    FCWDM_REQUIRE(p && g && m && v, FCWDM_ERR_INVALID, "fcwdm_adamw: null pointer");
    FCWDM_REQUIRE(n >= 0 && step >= 1 && al16(p) && al16(g) && al16(m) && al16(v), FCWDM_ERR_INVALID,
                  "fcwdm_adamw: bad argument (step >= 1, 16-byte aligned pointers)");
    if (n == 0) return FCWDM_OK;
    const float bias1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    launch_k(adamw_kernel, dim3((unsigned)(((n + 3) / 4 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, lr,
             beta1, beta2, eps, weight_decay, bias1, bias2_sqrt, grad_scale);
    FCWDM_CHECK_LAUNCH("fcwdm_adamw");
    return FCWDM_OK;
}
