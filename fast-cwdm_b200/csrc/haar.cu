// K1 / K2: 3-D Haar DWT / IDWT as a single 2x2x2 butterfly pass (HBM-bound).
//
// Replaces the reference's 14 band-matrix matmuls per transform (DWT_IDWT/DWT_IDWT_Functions.py:115-208)
// and the per-call rebuild + upload of six dense band matrices (DWT_IDWT/DWT_IDWT_layer.py:459-518).
// Algorithmic traffic: every input element read once, every output element written once.
//
// Planar (NCDHW) fast path: one thread owns a 2(D) x 2(H) x 8(W) brick: four 256-bit loads (fp32) or four
// 128-bit loads (bf16), eight 128-bit / 64-bit stores (one per band), all fully coalesced along W.
// Channels-last (NDHWC bf16) path used inside the denoiser: one thread owns 8 channels of a 2x2x2 brick.
#include "common.cuh"

namespace fcwdm {

__device__ __forceinline__ float apply_scale(float v, float scale) {
    // th.cat([LLL / 3., ...]) in the reference is a true division; keep it bit-identical for that case.
    return (scale == (1.0f / 3.0f)) ? (v / 3.0f) : (v * scale);
}

struct PlanarIdx {
    int64_t n, c, dd, hh, wq;
};
__device__ __forceinline__ PlanarIdx decompose(int64_t idx, int64_t C, int64_t D2, int64_t H2, int64_t WQ) {
    PlanarIdx r;
    r.wq = idx % WQ;
    int64_t t = idx / WQ;
    r.hh = t % H2;
    t /= H2;
    r.dd = t % D2;
    t /= D2;
    r.c = t % C;
    r.n = t / C;
    return r;
}

// ---------------------------------------------------------------------------------------------------
// planar fp32, W % 8 == 0
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwt3d_planar_f32_v8(const float* __restrict__ x, float* __restrict__ out,
                                                           int64_t total, int64_t C, int64_t D, int64_t H,
                                                           int64_t W, int64_t x_sn, int64_t x_sc, int64_t o_sn,
                                                           int64_t o_sc, int64_t o_sb, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    PlanarIdx p = decompose(idx, C, D2, H2, W >> 3);
    const float* src = x + p.n * x_sn + p.c * x_sc + ((2 * p.dd) * H + 2 * p.hh) * W + 8 * p.wq;
    float8 r[4];
    r[0] = ld_stream_f8(src);
    r[1] = ld_stream_f8(src + W);
    r[2] = ld_stream_f8(src + H * W);
    r[3] = ld_stream_f8(src + H * W + W);
    float ob[8][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float xin[8], b[8];
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
            xin[ij * 2 + 0] = r[ij].v[2 * q];
            xin[ij * 2 + 1] = r[ij].v[2 * q + 1];
        }
        haar_analysis(xin, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) ob[k][q] = b[k];
    }
    float* dst = out + p.n * o_sn + p.c * o_sc + (p.dd * H2 + p.hh) * W2 + 4 * p.wq;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float4 v = make_float4(ob[k][0], ob[k][1], ob[k][2], ob[k][3]);
        if (k == 0) {
            v.x = apply_scale(v.x, lll_scale);
            v.y = apply_scale(v.y, lll_scale);
            v.z = apply_scale(v.z, lll_scale);
            v.w = apply_scale(v.w, lll_scale);
        }
        st_stream_f4(dst + k * o_sb, v);
    }
}

__global__ void __launch_bounds__(256) idwt3d_planar_f32_v8(const float* __restrict__ bands, float* __restrict__ y,
                                                            int64_t total, int64_t C, int64_t D, int64_t H,
                                                            int64_t W, int64_t b_sn, int64_t b_sc, int64_t b_sb,
                                                            int64_t y_sn, int64_t y_sc, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    PlanarIdx p = decompose(idx, C, D2, H2, W >> 3);
    const float* src = bands + p.n * b_sn + p.c * b_sc + (p.dd * H2 + p.hh) * W2 + 4 * p.wq;
    float4 bv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bv[k] = ld_stream_f4(src + k * b_sb);
    bv[0].x *= lll_scale; bv[0].y *= lll_scale; bv[0].z *= lll_scale; bv[0].w *= lll_scale;
    float8 r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float b[8], xo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = (q == 0) ? bv[k].x : (q == 1) ? bv[k].y : (q == 2) ? bv[k].z : bv[k].w;
        haar_synthesis(b, xo);
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
            r[ij].v[2 * q] = xo[ij * 2];
            r[ij].v[2 * q + 1] = xo[ij * 2 + 1];
        }
    }
    float* dst = y + p.n * y_sn + p.c * y_sc + ((2 * p.dd) * H + 2 * p.hh) * W + 8 * p.wq;
    st_stream_f8(dst, r[0]);
    st_stream_f8(dst + W, r[1]);
    st_stream_f8(dst + H * W, r[2]);
    st_stream_f8(dst + H * W + W, r[3]);
}

// ---------------------------------------------------------------------------------------------------
// planar bf16, W % 8 == 0 (fp32 math)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwt3d_planar_bf16_v8(const __nv_bfloat16* __restrict__ x,
                                                            __nv_bfloat16* __restrict__ out, int64_t total,
                                                            int64_t C, int64_t D, int64_t H, int64_t W,
                                                            int64_t x_sn, int64_t x_sc, int64_t o_sn, int64_t o_sc,
                                                            int64_t o_sb, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    PlanarIdx p = decompose(idx, C, D2, H2, W >> 3);
    const __nv_bfloat16* src = x + p.n * x_sn + p.c * x_sc + ((2 * p.dd) * H + 2 * p.hh) * W + 8 * p.wq;
    float r[4][8];
    unpack8(ld_stream_u4(src), r[0]);
    unpack8(ld_stream_u4(src + W), r[1]);
    unpack8(ld_stream_u4(src + H * W), r[2]);
    unpack8(ld_stream_u4(src + H * W + W), r[3]);
    float ob[8][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float xin[8], b[8];
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
            xin[ij * 2 + 0] = r[ij][2 * q];
            xin[ij * 2 + 1] = r[ij][2 * q + 1];
        }
        haar_analysis(xin, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) ob[k][q] = b[k];
    }
    __nv_bfloat16* dst = out + p.n * o_sn + p.c * o_sc + (p.dd * H2 + p.hh) * W2 + 4 * p.wq;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float s = (k == 0) ? lll_scale : 1.0f;
        st_stream_u2(dst + k * o_sb, make_uint2(pack_bf16(ob[k][0] * s, ob[k][1] * s), pack_bf16(ob[k][2] * s, ob[k][3] * s)));
    }
}

__global__ void __launch_bounds__(256) idwt3d_planar_bf16_v8(const __nv_bfloat16* __restrict__ bands,
                                                             __nv_bfloat16* __restrict__ y, int64_t total,
                                                             int64_t C, int64_t D, int64_t H, int64_t W,
                                                             int64_t b_sn, int64_t b_sc, int64_t b_sb, int64_t y_sn,
                                                             int64_t y_sc, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    PlanarIdx p = decompose(idx, C, D2, H2, W >> 3);
    const __nv_bfloat16* src = bands + p.n * b_sn + p.c * b_sc + (p.dd * H2 + p.hh) * W2 + 4 * p.wq;
    float bv[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint2 u = ld_stream_u2(src + k * b_sb);
        bv[k][0] = bf16lo(u.x); bv[k][1] = bf16hi(u.x); bv[k][2] = bf16lo(u.y); bv[k][3] = bf16hi(u.y);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) bv[0][q] *= lll_scale;
    float r[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float b[8], xo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = bv[k][q];
        haar_synthesis(b, xo);
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
            r[ij][2 * q] = xo[ij * 2];
            r[ij][2 * q + 1] = xo[ij * 2 + 1];
        }
    }
    __nv_bfloat16* dst = y + p.n * y_sn + p.c * y_sc + ((2 * p.dd) * H + 2 * p.hh) * W + 8 * p.wq;
    st_stream_u4(dst, pack8(r[0]));
    st_stream_u4(dst + W, pack8(r[1]));
    st_stream_u4(dst + H * W, pack8(r[2]));
    st_stream_u4(dst + H * W + W, pack8(r[3]));
}

// ---------------------------------------------------------------------------------------------------
// planar generic (any even dims / any alignment): one thread per output voxel
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) dwt3d_planar_generic(const T* __restrict__ x, T* __restrict__ out,
                                                            int64_t total, int64_t C, int64_t D, int64_t H,
                                                            int64_t W, int64_t x_sn, int64_t x_sc, int64_t o_sn,
                                                            int64_t o_sc, int64_t o_sb, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    PlanarIdx p = decompose(idx, C, D2, H2, W2);
    const T* src = x + p.n * x_sn + p.c * x_sc + ((2 * p.dd) * H + 2 * p.hh) * W + 2 * p.wq;
    float xin[8], b[8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) xin[i * 4 + j * 2 + k] = to_f<T>(src[(i * H + j) * W + k]);
    haar_analysis(xin, b);
    b[0] = apply_scale(b[0], lll_scale);
    T* dst = out + p.n * o_sn + p.c * o_sc + (p.dd * H2 + p.hh) * W2 + p.wq;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k * o_sb] = from_f<T>(b[k]);
}

template <typename T>
__global__ void __launch_bounds__(256) idwt3d_planar_generic(const T* __restrict__ bands, T* __restrict__ y,
                                                             int64_t total, int64_t C, int64_t D, int64_t H,
                                                             int64_t W, int64_t b_sn, int64_t b_sc, int64_t b_sb,
                                                             int64_t y_sn, int64_t y_sc, float lll_scale) {
    pdl_prologue();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    PlanarIdx p = decompose(idx, C, D2, H2, W2);
    const T* src = bands + p.n * b_sn + p.c * b_sc + (p.dd * H2 + p.hh) * W2 + p.wq;
    float b[8], xo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = to_f<T>(src[k * b_sb]);
    b[0] *= lll_scale;
    haar_synthesis(b, xo);
    T* dst = y + p.n * y_sn + p.c * y_sc + ((2 * p.dd) * H + 2 * p.hh) * W + 2 * p.wq;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) dst[(i * H + j) * W + k] = from_f<T>(xo[i * 4 + j * 2 + k]);
}

// ---------------------------------------------------------------------------------------------------
// channels-last bf16 (denoiser-internal): one thread = 8 channels of one 2x2x2 brick
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwt3d_cl_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld,
                                                       __nv_bfloat16* __restrict__ lll, int64_t lll_ld,
                                                       __nv_bfloat16* __restrict__ hi, int64_t hi_ld, int64_t hi_sb,
                                                       const float* __restrict__ lll_bias, int64_t bias_ld,
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
    const __nv_bfloat16* src = x + (((n * D + 2 * dd) * H + 2 * hh) * W + 2 * ww) * x_ld + cq * 8;
    float in[8][8];  // [brick position i*4+j*2+k][channel]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                unpack8(ld_stream_u4(src + ((i * H + j) * W + k) * x_ld), in[i * 4 + j * 2 + k]);
    float ob[8][8];  // [band][channel]
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
    {
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = apply_scale(ob[0][c], lll_scale);
        if (lll_bias != nullptr) {
            const float4 b0 = *reinterpret_cast<const float4*>(lll_bias + n * bias_ld + cq * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(lll_bias + n * bias_ld + cq * 8 + 4);
            o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
            o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
        }
        *reinterpret_cast<uint4*>(lll + vox * lll_ld + cq * 8) = pack8(o);
    }
    if (hi != nullptr) {
#pragma unroll
        for (int b = 1; b < 8; ++b) {
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = apply_scale(ob[b][c], hi_scale);
            *reinterpret_cast<uint4*>(hi + (b - 1) * hi_sb + vox * hi_ld + cq * 8) = pack8(o);
        }
    }
}

__global__ void __launch_bounds__(256) idwt3d_cl_kernel(const __nv_bfloat16* __restrict__ lll, int64_t lll_ld,
                                                        const __nv_bfloat16* __restrict__ hi, int64_t hi_ld,
                                                        int64_t hi_sb, __nv_bfloat16* __restrict__ y, int64_t y_ld,
                                                        const float* __restrict__ bias, int64_t bias_ld,
                                                        int64_t total, int64_t D, int64_t H, int64_t W, int64_t C,
                                                        float lll_scale) {
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
    float bnd[8][8];  // [band][channel]
    unpack8(ld_stream_u4(lll + vox * lll_ld + cq * 8), bnd[0]);
#pragma unroll
    for (int c = 0; c < 8; ++c) bnd[0][c] *= lll_scale;
#pragma unroll
    for (int b = 1; b < 8; ++b) unpack8(ld_stream_u4(hi + (b - 1) * hi_sb + vox * hi_ld + cq * 8), bnd[b]);
    float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + n * bias_ld + cq * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + n * bias_ld + cq * 8 + 4);
        bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w;
        bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
    }
    float out[8][8];  // [brick position][channel]
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float b[8], xo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] = bnd[q][c];
        haar_synthesis(b, xo);
#pragma unroll
        for (int q = 0; q < 8; ++q) out[q][c] = xo[q] + bs[c];
    }
    __nv_bfloat16* dst = y + (((n * D + 2 * dd) * H + 2 * hh) * W + 2 * ww) * y_ld + cq * 8;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                *reinterpret_cast<uint4*>(dst + ((i * H + j) * W + k) * y_ld) = pack8(out[i * 4 + j * 2 + k]);
}

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static int check_dims(const char* fn, const void* a, const void* b, int64_t N, int64_t C, int64_t D, int64_t H,
                      int64_t W) {
    FCWDM_REQUIRE(N >= 0 && C >= 0 && D >= 0 && H >= 0 && W >= 0, FCWDM_ERR_INVALID, "%s: negative dimension", fn);
    FCWDM_REQUIRE((D % 2 == 0) && (H % 2 == 0) && (W % 2 == 0), FCWDM_ERR_UNSUPPORTED,
                  "%s: D, H, W must be even (got %lld, %lld, %lld)", fn, (long long)D, (long long)H, (long long)W);
    // empty tensors legitimately carry null data pointers
    FCWDM_REQUIRE((a != nullptr && b != nullptr) || (N * C * D * H * W == 0), FCWDM_ERR_INVALID, "%s: null pointer", fn);
    return FCWDM_OK;
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_dwt3d_fwd(const void* x, void* bands, int dtype, int64_t N, int64_t C, int64_t D, int64_t H,
                               int64_t W, int64_t x_sn, int64_t x_sc, int64_t o_sn, int64_t o_sc, int64_t o_sb,
                               float lll_scale, void* stream) {
    int rc = check_dims("fcwdm_dwt3d_fwd", x, bands, N, C, D, H, W);
    if (rc) return rc;
    FCWDM_REQUIRE(dtype == FCWDM_F32 || dtype == FCWDM_BF16, FCWDM_ERR_INVALID, "fcwdm_dwt3d_fwd: bad dtype %d", dtype);
    if (N * C * D * H * W == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const bool strides_v8 = (W % 8 == 0) && (x_sn % 8 == 0) && (x_sc % 8 == 0) && (o_sn % 4 == 0) && (o_sc % 4 == 0) &&
                            (o_sb % 4 == 0);
    if (dtype == FCWDM_F32) {
        if (strides_v8 && aligned(x, 32) && aligned(bands, 16)) {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 8);
            launch_k(dwt3d_planar_f32_v8, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const float*)x, (float*)bands, total, C, D, H, W, x_sn, x_sc, o_sn, o_sc, o_sb, lll_scale);
        } else {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 2);
            launch_k(dwt3d_planar_generic<float>, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const float*)x, (float*)bands, total, C, D, H, W, x_sn, x_sc, o_sn, o_sc, o_sb, lll_scale);
        }
    } else {
        if (strides_v8 && aligned(x, 16) && aligned(bands, 8)) {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 8);
            launch_k(dwt3d_planar_bf16_v8, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const __nv_bfloat16*)x, (__nv_bfloat16*)bands, total, C, D, H, W, x_sn, x_sc, o_sn, o_sc, o_sb,
                lll_scale);
        } else {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 2);
            launch_k(dwt3d_planar_generic<__nv_bfloat16>, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const __nv_bfloat16*)x, (__nv_bfloat16*)bands, total, C, D, H, W, x_sn, x_sc, o_sn, o_sc, o_sb,
                lll_scale);
        }
    }
    FCWDM_CHECK_LAUNCH("fcwdm_dwt3d_fwd");
    return FCWDM_OK;
}

extern "C" int fcwdm_idwt3d_fwd(const void* bands, void* y, int dtype, int64_t N, int64_t C, int64_t D, int64_t H,
                                int64_t W, int64_t b_sn, int64_t b_sc, int64_t b_sb, int64_t y_sn, int64_t y_sc,
                                float lll_scale, void* stream) {
    int rc = check_dims("fcwdm_idwt3d_fwd", bands, y, N, C, D, H, W);
    if (rc) return rc;
    FCWDM_REQUIRE(dtype == FCWDM_F32 || dtype == FCWDM_BF16, FCWDM_ERR_INVALID, "fcwdm_idwt3d_fwd: bad dtype %d", dtype);
    if (N * C * D * H * W == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const bool strides_v8 = (W % 8 == 0) && (y_sn % 8 == 0) && (y_sc % 8 == 0) && (b_sn % 4 == 0) && (b_sc % 4 == 0) &&
                            (b_sb % 4 == 0);
    if (dtype == FCWDM_F32) {
        if (strides_v8 && aligned(y, 32) && aligned(bands, 16)) {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 8);
            launch_k(idwt3d_planar_f32_v8, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const float*)bands, (float*)y, total, C, D, H, W, b_sn, b_sc, b_sb, y_sn, y_sc, lll_scale);
        } else {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 2);
            launch_k(idwt3d_planar_generic<float>, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const float*)bands, (float*)y, total, C, D, H, W, b_sn, b_sc, b_sb, y_sn, y_sc, lll_scale);
        }
    } else {
        if (strides_v8 && aligned(y, 16) && aligned(bands, 8)) {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 8);
            launch_k(idwt3d_planar_bf16_v8, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const __nv_bfloat16*)bands, (__nv_bfloat16*)y, total, C, D, H, W, b_sn, b_sc, b_sb, y_sn, y_sc,
                lll_scale);
        } else {
            int64_t total = N * C * (D / 2) * (H / 2) * (W / 2);
            launch_k(idwt3d_planar_generic<__nv_bfloat16>, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, st, 
                (const __nv_bfloat16*)bands, (__nv_bfloat16*)y, total, C, D, H, W, b_sn, b_sc, b_sb, y_sn, y_sc,
                lll_scale);
        }
    }
    FCWDM_CHECK_LAUNCH("fcwdm_idwt3d_fwd");
    return FCWDM_OK;
}

extern "C" int fcwdm_dwt3d_cl(const void* x, int64_t x_ld, void* lll, int64_t lll_ld, void* hi, int64_t hi_ld,
                              int64_t hi_sb, const float* lll_bias, int64_t bias_ld, int64_t N, int64_t D, int64_t H,
                              int64_t W, int64_t C, float lll_scale, float hi_scale, void* stream) {
    int rc = check_dims("fcwdm_dwt3d_cl", x, lll, N, C, D, H, W);
    if (rc) return rc;
    FCWDM_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && lll_ld % 8 == 0 && (hi == nullptr || (hi_ld % 8 == 0 && hi_sb % 8 == 0)),
                  FCWDM_ERR_UNSUPPORTED, "fcwdm_dwt3d_cl: C and strides must be multiples of 8");
    FCWDM_REQUIRE(aligned(x, 16) && aligned(lll, 16) && aligned(hi, 16) && aligned(lll_bias, 16) && bias_ld % 4 == 0,
                  FCWDM_ERR_INVALID, "fcwdm_dwt3d_cl: pointers must be 16-byte aligned");
    int64_t total = N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(dwt3d_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)lll, lll_ld, (__nv_bfloat16*)hi, hi_ld, hi_sb, lll_bias, bias_ld,
        total, D, H, W, C, lll_scale, hi_scale);
    FCWDM_CHECK_LAUNCH("fcwdm_dwt3d_cl");
    return FCWDM_OK;
}

extern "C" int fcwdm_idwt3d_cl(const void* lll, int64_t lll_ld, const void* hi, int64_t hi_ld, int64_t hi_sb, void* y,
                               int64_t y_ld, const float* bias, int64_t bias_ld, int64_t N, int64_t D, int64_t H, int64_t W,
                               int64_t C, float lll_scale, void* stream) {
    int rc = check_dims("fcwdm_idwt3d_cl", lll, y, N, C, D, H, W);
    if (rc) return rc;
    FCWDM_REQUIRE(hi != nullptr, FCWDM_ERR_INVALID, "fcwdm_idwt3d_cl: null high-band pointer");
    FCWDM_REQUIRE(C % 8 == 0 && y_ld % 8 == 0 && lll_ld % 8 == 0 && hi_ld % 8 == 0 && hi_sb % 8 == 0,
                  FCWDM_ERR_UNSUPPORTED, "fcwdm_idwt3d_cl: C and strides must be multiples of 8");
    FCWDM_REQUIRE(aligned(y, 16) && aligned(lll, 16) && aligned(hi, 16) && aligned(bias, 16) && bias_ld % 4 == 0,
                  FCWDM_ERR_INVALID, "fcwdm_idwt3d_cl: pointers must be 16-byte aligned");
    int64_t total = N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(idwt3d_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)lll, lll_ld, (const __nv_bfloat16*)hi, hi_ld, hi_sb, (__nv_bfloat16*)y, y_ld, bias, bias_ld,
        total, D, H, W, C, lll_scale);
    FCWDM_CHECK_LAUNCH("fcwdm_idwt3d_cl");
    return FCWDM_OK;
}
