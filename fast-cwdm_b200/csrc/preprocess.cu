// BraTS volume preprocessing on the GPU (guided_diffusion/bratsloader.py:44-50,107-111): per-volume quantile clip
// (np.quantile 0.1 % / 99.9 %, linear interpolation), min-max normalisation to [0, 1], zero-pad the slice axis
// 155 -> 160 and crop 240 x 240 -> 224 x 224.  The reference does this on the host in float64 numpy inside the
// DataLoader workers (a full sort-based quantile of 8.9 M voxels, four times per case); here raw volumes go
// disk -> GPU once and everything is HBM-bound passes:
//   * exact order statistics by MSB-first radix select on the order-preserving integer image of the fp32 bits:
//     four 8-bit histogram passes (shared-memory histograms, one global atomic per bin and block), each pass
//     narrowing the prefix of all requested ranks of all volumes at once;
//   * one elementwise pass clip -> normalise -> pad -> crop.
#include <math.h>

#include "common.cuh"

namespace fcwdm {

constexpr int kRanks = 4;     // floor/ceil positions of the two quantiles

__device__ __forceinline__ uint32_t f32_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct SelectState {            // per (volume, rank)
    unsigned long long rank;    // remaining rank inside the current prefix bucket
    uint32_t prefix;            // key bits fixed so far (high bits)
    uint32_t pad;
};

// pass p (0 = most significant byte): every block first folds the previous pass' histogram into (prefix, rank) -- 256
// bins, done redundantly in the prologue -- then histograms byte p of the elements matching the prefix.  The state is
// double-buffered: pass p reads state[p & 1] (folded through pass p-2) and block 0 writes the folded state to
// state[(p+1) & 1], so no block can read a state another block of the same launch has already advanced.
__global__ void __launch_bounds__(256) select_pass_kernel(const float* __restrict__ x, int64_t n_per, int pass, int n_vr,
                                                          SelectState* __restrict__ state /* [2][V*kRanks] */,
                                                          unsigned int* __restrict__ hist /* [V*kRanks][4][256] */) {
    pdl_prologue();
    __shared__ unsigned int sh[256];
    __shared__ uint32_t s_prefix;
    const int vr = blockIdx.y;                  // volume * kRanks + rank index
    const int vol = vr / kRanks;
    if (threadIdx.x == 0) {
        SelectState st = state[(size_t)(pass & 1) * n_vr + vr];
        if (pass > 0) {
            const unsigned int* h = hist + ((size_t)vr * 4 + (pass - 1)) * 256;
            unsigned long long r = st.rank;
            int b = 0;
            for (; b < 255; ++b) {
                const unsigned int c = h[b];
                if (r < c) break;
                r -= c;
            }
            st.rank = r;
            st.prefix |= (uint32_t)b << (8 * (4 - pass));
        }
        s_prefix = st.prefix;
        if (blockIdx.x == 0) state[(size_t)((pass + 1) & 1) * n_vr + vr] = st;
    }
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const int shift = 8 * (3 - pass);
    const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    const float* xv = x + (size_t)vol * n_per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = f32_key(xv[i]);
        if ((k & mask) == (prefix & mask)) atomicAdd(&sh[(k >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist + ((size_t)vr * 4 + pass) * 256 + threadIdx.x, sh[threadIdx.x]);
}

// fold the last histogram, produce the two interpolated quantiles per volume (numpy's 'linear' method, float64)
__global__ void select_finish_kernel(SelectState* __restrict__ state, const unsigned int* __restrict__ hist,
                                     const double* __restrict__ frac /* [2] */, float* __restrict__ q_out /* [V][2] */,
                                     int V) {
    pdl_prologue();
    const int vol = blockIdx.x * blockDim.x + threadIdx.x;
    if (vol >= V) return;
    double val[kRanks];
    for (int r = 0; r < kRanks; ++r) {
        const int vr = vol * kRanks + r;
        SelectState st = state[vr];                  // buffer 0 = folded through pass 2 (written by pass 3)
        const unsigned int* h = hist + ((size_t)vr * 4 + 3) * 256;
        unsigned long long rem = st.rank;
        int b = 0;
        for (; b < 255; ++b) {
            const unsigned int c = h[b];
            if (rem < c) break;
            rem -= c;
        }
        val[r] = (double)key_f32(st.prefix | (uint32_t)b);
    }
    for (int q = 0; q < 2; ++q) {
        const double a = val[2 * q], b = val[2 * q + 1], t = frac[q];
        // numpy _lerp: a + (b - a) * t, evaluated from the b side for t >= 0.5
        double v = a + (b - a) * t;
        if (t >= 0.5) v = b - (b - a) * (1.0 - t);
        if (t == 0.0) v = a;
        q_out[vol * 2 + q] = (float)v;
    }
}

// out (V, 1, X - 2 cx, Y - 2 cy, Zp) = pad_z(crop_xy((clip(x, qlo, qhi) - qlo) / (qhi - qlo)))
__global__ void __launch_bounds__(256) clipnorm_kernel(const float* __restrict__ x, const float* __restrict__ q,
                                                       float* __restrict__ out, int64_t total, int X, int Y, int Z, int cx,
                                                       int cy, int Zp) {
    pdl_prologue();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Xo = X - 2 * cx, Yo = Y - 2 * cy;
    const int z = (int)(idx % Zp);
    int64_t t = idx / Zp;
    const int yy = (int)(t % Yo); t /= Yo;
    const int xx = (int)(t % Xo);
    const int vol = (int)(t / Xo);
    float v = 0.f;
    if (z < Z) {
        const float lo = q[vol * 2 + 0], hi = q[vol * 2 + 1];
        const float raw = x[(((size_t)vol * X + xx + cx) * Y + yy + cy) * Z + z];
        const float c = fminf(fmaxf(raw, lo), hi);
        v = (c - lo) / (hi - lo);          // constant volume: 0/0 = NaN, as the reference's numpy expression
    }
    out[idx] = v;
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int64_t fcwdm_clip_normalize_workspace_bytes(int64_t V) {
    if (V < 0) return -1;
    return V * kRanks * (int64_t)(2 * sizeof(SelectState) + 4 * 256 * sizeof(unsigned int)) + 64;
}

extern "C" int fcwdm_clip_normalize(const float* x, float* out, float* quantiles, void* workspace, int64_t workspace_bytes,
                                    int64_t V, int64_t X, int64_t Y, int64_t Z, int64_t crop_x, int64_t crop_y,
                                    int64_t pad_z_to, double q_lo, double q_hi, void* stream) {
    FCWDM_REQUIRE(x && out && quantiles && workspace, FCWDM_ERR_INVALID, "fcwdm_clip_normalize: null pointer");
    FCWDM_REQUIRE(V >= 0 && X > 0 && Y > 0 && Z > 0 && crop_x >= 0 && crop_y >= 0 && 2 * crop_x < X && 2 * crop_y < Y &&
                      pad_z_to >= Z && q_lo >= 0.0 && q_lo <= q_hi && q_hi <= 1.0 && V <= 16383,
                  FCWDM_ERR_INVALID, "fcwdm_clip_normalize: bad argument");
    FCWDM_REQUIRE(workspace_bytes >= fcwdm_clip_normalize_workspace_bytes(V), FCWDM_ERR_INVALID,
                  "fcwdm_clip_normalize: workspace of %lld bytes needed", (long long)fcwdm_clip_normalize_workspace_bytes(V));
    if (V == 0) return FCWDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_per = X * Y * Z;
    // workspace layout: frac[2] doubles (64 B) | SelectState[2][V*4] | hist[V*4][4][256]
    double* frac = (double*)workspace;
    SelectState* state = (SelectState*)((char*)workspace + 64);
    unsigned int* hist = (unsigned int*)((char*)state + 2 * V * kRanks * sizeof(SelectState));
    // ranks: numpy 'linear': virtual index q * (n - 1); floor / ceil neighbours
    SelectState h_state[kRanks];
    double h_frac[2];
    const double qs[2] = {q_lo, q_hi};
    for (int q = 0; q < 2; ++q) {
        const double pos = qs[q] * (double)(n_per - 1);
        const double fl = floor(pos);
        h_frac[q] = pos - fl;
        const unsigned long long lo = (unsigned long long)fl;
        const unsigned long long hi = lo + 1 < (unsigned long long)n_per ? lo + 1 : lo;
        h_state[2 * q] = SelectState{lo, 0u, 0u};
        h_state[2 * q + 1] = SelectState{hi, 0u, 0u};
    }
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)V * kRanks * 4 * 256 * sizeof(unsigned int), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(frac, h_frac, sizeof(h_frac), cudaMemcpyHostToDevice, st);
    for (int64_t v = 0; v < V && e == cudaSuccess; ++v)
        e = cudaMemcpyAsync(state + v * kRanks, h_state, sizeof(h_state), cudaMemcpyHostToDevice, st);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_clip_normalize: workspace setup failed (%s)", cudaGetErrorString(e));
    int bx = (int)((n_per + 256 * 16 - 1) / (256 * 16));
    const int cap = num_sms() * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    for (int pass = 0; pass < 4; ++pass) {
        launch_k(select_pass_kernel, dim3((unsigned)bx, (unsigned)(V * kRanks)), dim3(256), 0, st, x, n_per, pass,
                 (int)(V * kRanks), state, hist);
        FCWDM_CHECK_LAUNCH("fcwdm_clip_normalize (select)");
    }
    launch_k(select_finish_kernel, dim3((unsigned)((V + 63) / 64)), dim3(64), 0, st, state, (const unsigned int*)hist,
             (const double*)frac, quantiles, (int)V);
    FCWDM_CHECK_LAUNCH("fcwdm_clip_normalize (finish)");
    const int64_t total = V * (X - 2 * crop_x) * (Y - 2 * crop_y) * pad_z_to;
    launch_k(clipnorm_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, (const float*)quantiles, out, total,
             (int)X, (int)Y, (int)Z, (int)crop_x, (int)crop_y, (int)pad_z_to);
    FCWDM_CHECK_LAUNCH("fcwdm_clip_normalize (apply)");
    return FCWDM_OK;
}
