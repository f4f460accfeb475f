// Shared helpers for the fcwdm sm_100a kernels (C-ABI declared in include/fcwdm.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fcwdm.h"

namespace fcwdm {

// Thread-local last-error text, returned by fcwdm_last_error().  The C-ABI never throws.
void set_error(const char* fmt, ...);

#define FCWDM_REQUIRE(cond, code, ...)                  \
    do {                                                \
        if (!(cond)) {                                  \
            ::fcwdm::set_error(__VA_ARGS__);            \
            return (code);                              \
        }                                               \
    } while (0)

#define FCWDM_CHECK_LAUNCH(name)                                                        \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            ::fcwdm::set_error("%s: CUDA error %d (%s)", name, (int)e__,               \
                               cudaGetErrorString(e__));                                \
            return FCWDM_ERR_CUDA;                                                      \
        }                                                                               \
    } while (0)

static inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Programmatic dependent launch: every kernel is launched with programmaticStreamSerialization so that its
// prologue (and its launch latency) overlaps the tail of its predecessor in the stream / captured graph; every
// kernel calls pdl_prologue() before its first global-memory access, which waits for the full completion (and
// memory flush) of the predecessor, so ordering semantics are unchanged.  FCWDM_NO_PDL=1 disables the attribute.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// the same with a runtime thread-block cluster of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct __align__(32) float8 {
    float v[8];
};

// 256-bit streaming global load / store (LDG.E.256 / STG.E.256 on sm_100a).
__device__ __forceinline__ float8 ld_stream_f8(const float* p) {
    float8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]),
                   "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f8(float* p, const float8& r) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]),
                 "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 r) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.x), "f"(r.y), "f"(r.z),
                 "f"(r.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 r) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z),
                 "r"(r.w)
                 : "memory");
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u2(void* p, uint2 r) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(r.x), "r"(r.y) : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 8 bf16 (uint4) <-> 8 floats
__device__ __forceinline__ void unpack8(uint4 u, float* f) {
    f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
    f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// Haar 2x2x2 butterfly on one brick.  x is indexed [i*4 + j*2 + k] for offsets (i,j,k) along (D,H,W);
// b is indexed by band = fd*4 + fh*2 + fw (LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH).
// Stage order follows DWTFunction_3D.forward (DWT_IDWT_Functions.py:122-135): H, then W, then D,
// each stage low = s*a + s*b, high = s*a - s*b in fp32.
__device__ __forceinline__ void haar_analysis(const float* x, float* b) {
    const float s = 0.70710678118654752440f;
    float t1[8], t2[8];
    // along H (j): pairs (i, 0, k) / (i, 1, k)  -> index [i*4 + fh*2 + k]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float a = x[i * 4 + 0 + k], c = x[i * 4 + 2 + k];
            t1[i * 4 + 0 + k] = s * a + s * c;
            t1[i * 4 + 2 + k] = s * a - s * c;
        }
    // along W (k) -> index [i*4 + fh*2 + fw]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float a = t1[i * 4 + j * 2 + 0], c = t1[i * 4 + j * 2 + 1];
            t2[i * 4 + j * 2 + 0] = s * a + s * c;
            t2[i * 4 + j * 2 + 1] = s * a - s * c;
        }
    // along D (i) -> band [fd*4 + fh*2 + fw]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float a = t2[r], c = t2[4 + r];
        b[r] = s * a + s * c;
        b[4 + r] = s * a - s * c;
    }
}

// Inverse, stage order of IDWTFunction_3D.forward (DWT_IDWT_Functions.py:167-180): D, then W, then H.
__device__ __forceinline__ void haar_synthesis(const float* b, float* x) {
    const float s = 0.70710678118654752440f;
    float t1[8], t2[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {  // D: (band r, band 4+r) -> i = 0/1
        float lo = b[r], hi = b[4 + r];
        t1[r] = s * lo + s * hi;
        t1[4 + r] = s * lo - s * hi;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // W: fw=0/1 -> k = 0/1
            float lo = t1[i * 4 + j * 2 + 0], hi = t1[i * 4 + j * 2 + 1];
            t2[i * 4 + j * 2 + 0] = s * lo + s * hi;
            t2[i * 4 + j * 2 + 1] = s * lo - s * hi;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // H: fh=0/1 -> j = 0/1
            float lo = t2[i * 4 + 0 + k], hi = t2[i * 4 + 2 + k];
            x[i * 4 + 0 + k] = s * lo + s * hi;
            x[i * 4 + 2 + k] = s * lo - s * hi;
        }
}

}  // namespace fcwdm
