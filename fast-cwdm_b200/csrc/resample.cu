// Average-pool / nearest-neighbour resampling on channels-last bf16 activations: the up/down-sampling of the plain
// UNetModel's ResBlocks (guided_diffusion/unet.py:40-100: F.interpolate(mode="nearest") and avg_pool_nd with
// kernel = stride = 2, or (1,2,2) when resample_2d).  HBM-bound, one thread = 8 channels of one OUTPUT voxel
// (pool) / one INPUT voxel (upsample: each loaded vector is stored fd*4 times).
#include "common.cuh"

namespace fcwdm {

// y = scale * (sum over the 2x2x(fd) brick of x) (+ acc): scale = 1/(4 fd) is the average pool, scale = 1 the adjoint of
// the nearest-neighbour up-sampling
__global__ void __launch_bounds__(256) avgpool2_cl_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld,
                                                          const __nv_bfloat16* __restrict__ acc_in, int64_t acc_ld,
                                                          __nv_bfloat16* __restrict__ y, int64_t y_ld, int64_t total,
                                                          int64_t D, int64_t H, int64_t W, int64_t C, int fd, float scale) {
    pdl_prologue();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t C8 = C >> 3, D2 = D / fd, H2 = H >> 1, W2 = W >> 1;
    const int64_t cq = idx % C8;
    int64_t t = idx / C8;
    const int64_t ww = t % W2; t /= W2;
    const int64_t hh = t % H2; t /= H2;
    const int64_t dd = t % D2;
    const int64_t n = t / D2;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < fd; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float f[8];
                unpack8(ld_stream_u4(x + (((n * D + dd * fd + i) * H + 2 * hh + j) * W + 2 * ww + k) * x_ld + cq * 8), f);
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] += f[c];
            }
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] *= scale;
    const int64_t vox = ((n * D2 + dd) * H2 + hh) * W2 + ww;
    if (acc_in != nullptr) {
        float fa[8];
        unpack8(ld_stream_u4(acc_in + vox * acc_ld + cq * 8), fa);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += fa[c];
    }
    *reinterpret_cast<uint4*>(y + vox * y_ld + cq * 8) = pack8(acc);
}

// y[brick] = scale * x (+ acc[brick]): scale = 1 is the nearest-neighbour up-sampling, scale = 1/(4 fd) the adjoint of the
// average pool
__global__ void __launch_bounds__(256) upsample2_cl_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld,
                                                           const __nv_bfloat16* __restrict__ acc_in, int64_t acc_ld,
                                                           __nv_bfloat16* __restrict__ y, int64_t y_ld, int64_t total,
                                                           int64_t D, int64_t H, int64_t W, int64_t C, int fd, float scale) {
    pdl_prologue();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t C8 = C >> 3;
    const int64_t cq = idx % C8;
    int64_t t = idx / C8;
    const int64_t ww = t % W; t /= W;
    const int64_t hh = t % H; t /= H;
    const int64_t dd = t % D;
    const int64_t n = t / D;
    uint4 v = ld_stream_u4(x + (((n * D + dd) * H + hh) * W + ww) * x_ld + cq * 8);
    float f[8];
    if (scale != 1.0f || acc_in != nullptr) {
        unpack8(v, f);
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] *= scale;
        v = pack8(f);
    }
    const int64_t D2 = D * fd, H2 = 2 * H, W2 = 2 * W;
    for (int i = 0; i < fd; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int64_t o = ((n * D2 + dd * fd + i) * H2 + 2 * hh + j) * W2 + 2 * ww + k;
                uint4 out = v;
                if (acc_in != nullptr) {
                    float fa[8], fo[8];
                    unpack8(ld_stream_u4(acc_in + o * acc_ld + cq * 8), fa);
#pragma unroll
                    for (int c = 0; c < 8; ++c) fo[c] = f[c] + fa[c];
                    out = pack8(fo);
                }
                *reinterpret_cast<uint4*>(y + o * y_ld + cq * 8) = out;
            }
}

static inline bool rs_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_avgpool2_cl(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t N, int64_t D, int64_t H,
                                 int64_t W, int64_t C, int pool_depth, void* stream) {
    FCWDM_REQUIRE(x && y, FCWDM_ERR_INVALID, "fcwdm_avgpool2_cl: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0, FCWDM_ERR_INVALID, "fcwdm_avgpool2_cl: bad dimension");
    // avg_pool3d floors odd sizes; the U-Net only ever halves even sizes, so odd sizes are refused instead
    FCWDM_REQUIRE(H % 2 == 0 && W % 2 == 0 && (!pool_depth || D % 2 == 0), FCWDM_ERR_UNSUPPORTED,
                  "fcwdm_avgpool2_cl: pooled dimensions must be even");
    FCWDM_REQUIRE(C % 8 == 0 && x_ld >= C && y_ld >= C && x_ld % 8 == 0 && y_ld % 8 == 0 && rs_al16(x) && rs_al16(y),
                  FCWDM_ERR_INVALID, "fcwdm_avgpool2_cl: C and strides must be multiples of 8, pointers 16-byte aligned");
    const int fd = pool_depth ? 2 : 1;
    const int64_t total = N * (D / fd) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(avgpool2_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)nullptr, (int64_t)0, (__nv_bfloat16*)y, y_ld, total, D, H, W, C, fd,
             1.0f / (float)(4 * fd));
    FCWDM_CHECK_LAUNCH("fcwdm_avgpool2_cl");
    return FCWDM_OK;
}

extern "C" int fcwdm_upsample2_cl(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t N, int64_t D, int64_t H,
                                  int64_t W, int64_t C, int up_depth, void* stream) {
    FCWDM_REQUIRE(x && y, FCWDM_ERR_INVALID, "fcwdm_upsample2_cl: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0, FCWDM_ERR_INVALID, "fcwdm_upsample2_cl: bad dimension");
    FCWDM_REQUIRE(C % 8 == 0 && x_ld >= C && y_ld >= C && x_ld % 8 == 0 && y_ld % 8 == 0 && rs_al16(x) && rs_al16(y),
                  FCWDM_ERR_INVALID, "fcwdm_upsample2_cl: C and strides must be multiples of 8, pointers 16-byte aligned");
    const int64_t total = N * D * H * W * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(upsample2_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)nullptr, (int64_t)0, (__nv_bfloat16*)y, y_ld, total, D, H, W, C,
             up_depth ? 2 : 1, 1.0f);
    FCWDM_CHECK_LAUNCH("fcwdm_upsample2_cl");
    return FCWDM_OK;
}

// adjoint of fcwdm_avgpool2_cl: dx[brick] = dy / (4 fd) (+ acc); (D,H,W) = dims of dy (the POOLED tensor)
extern "C" int fcwdm_avgpool2_cl_bwd(const void* dy, int64_t dy_ld, const void* acc, int64_t acc_ld, void* dx, int64_t dx_ld,
                                     int64_t N, int64_t D, int64_t H, int64_t W, int64_t C, int pool_depth, void* stream) {
    FCWDM_REQUIRE(dy && dx, FCWDM_ERR_INVALID, "fcwdm_avgpool2_cl_bwd: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0 && C % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0 &&
                      acc_ld % 8 == 0 && rs_al16(dy) && rs_al16(dx) && rs_al16(acc),
                  FCWDM_ERR_INVALID, "fcwdm_avgpool2_cl_bwd: bad argument");
    const int fd = pool_depth ? 2 : 1;
    const int64_t total = N * D * H * W * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(upsample2_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)acc, acc_ld, (__nv_bfloat16*)dx, dx_ld, total, D, H, W, C, fd,
             1.0f / (float)(4 * fd));
    FCWDM_CHECK_LAUNCH("fcwdm_avgpool2_cl_bwd");
    return FCWDM_OK;
}

// adjoint of fcwdm_upsample2_cl: dx = sum of dy over each brick (+ acc); (D,H,W) = dims of dy (the UP-SAMPLED tensor)
extern "C" int fcwdm_upsample2_cl_bwd(const void* dy, int64_t dy_ld, const void* acc, int64_t acc_ld, void* dx, int64_t dx_ld,
                                      int64_t N, int64_t D, int64_t H, int64_t W, int64_t C, int up_depth, void* stream) {
    FCWDM_REQUIRE(dy && dx, FCWDM_ERR_INVALID, "fcwdm_upsample2_cl_bwd: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0 && C % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0 &&
                      acc_ld % 8 == 0 && rs_al16(dy) && rs_al16(dx) && rs_al16(acc) && H % 2 == 0 && W % 2 == 0 &&
                      (!up_depth || D % 2 == 0),
                  FCWDM_ERR_INVALID, "fcwdm_upsample2_cl_bwd: bad argument");
    const int fd = up_depth ? 2 : 1;
    const int64_t total = N * (D / fd) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return FCWDM_OK;
    launch_k(avgpool2_cl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)acc, acc_ld, (__nv_bfloat16*)dx, dx_ld, total, D, H, W, C, fd,
             1.0f);
    FCWDM_CHECK_LAUNCH("fcwdm_upsample2_cl_bwd");
    return FCWDM_OK;
}
